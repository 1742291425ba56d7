"""CPU, world_size 2, gloo: the host-side sharding / gathering logic of parallel.py (the N>1 path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_reward(f_rew, t, q, offs, model, param):
    # deterministic stand-in for the GPU reward: per-path sum of a function of the points
    return np.array([np.sum(np.sin(q[offs[p]:offs[p + 1]]).sum(1) * (p + 1)) for p in range(len(offs) - 1)])


def _worker(rank, size, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    from informative_path_planning_b200 import parallel
    rng = np.random.default_rng(0)
    P = 11
    lens = rng.integers(1, 7, P)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    q = rng.uniform(0, 10, (offsets[-1], 2))
    got = parallel.reward_paths_sharded(None, 'mean', 0, q, offsets, reward_fn=_fake_reward)
    # every rank must hold the full, correctly ordered vector
    want = np.array([np.sum(np.sin(q[offsets[p]:offsets[p + 1]]).sum(1) * 1) for p in range(P)])
    plo, phi, qlo, qhi = parallel.shard_paths(offsets, rank, size)
    local_idx = np.arange(phi - plo) + 1
    full = []
    for r in range(size):
        a, b, _, _ = parallel.shard_paths(offsets, r, size)
        full.extend((np.arange(b - a) + 1).tolist())
    want = want * np.array(full)
    b = parallel.broadcast_array(np.arange(12.0).reshape(3, 4) if rank == 0 else None, src=0)
    g = parallel.gather_concat(torch.full((rank + 1,), float(rank), dtype=torch.float64), [1, 2])
    np.save(os.path.join(tmp, 'r%d.npy' % rank), np.concatenate([got, want, b.ravel(), g.numpy()]))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from informative_path_planning_b200.parallel import shard_bounds, shard_paths
    for n in (0, 1, 7, 8, 65536):
        for size in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(size - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    offs = np.array([0, 3, 3, 10, 12, 20])
    assert shard_paths(offs, 1, 2) == (3, 5, 10, 20)


def test_gloo_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / 'r0.npy')
    r1 = np.load(tmp_path / 'r1.npy')
    np.testing.assert_array_equal(r0, r1)
    P = 11
    np.testing.assert_allclose(r0[:P], r0[P:2 * P])
    np.testing.assert_array_equal(r0[2 * P:2 * P + 12], np.arange(12.0))
    np.testing.assert_array_equal(r0[2 * P + 12:], [0.0, 1.0, 1.0])


def _mcts_worker(rank, size, port, tmp):
    import contextlib
    import io
    import sys
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    from informative_path_planning_b200 import mcts_library as mc
    from informative_path_planning_b200 import parallel
    from informative_path_planning_b200.paths_library import Path_Generator
    from oracle import aq_oracle
    from oracle.gpmodel_oracle import OnlineGPModelOracle
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 10, (30, 2))
    z = np.sin(X[:, :1]) * 5 + rng.normal(0, 0.5, (30, 1))
    m = OnlineGPModelOracle((0., 10., 0., 10.), 1.0, 100.0, noise=0.5, dimension=2)      # the replicated belief
    m.add_data(X, z)
    pg = Path_Generator(8, 1.5, 0.05, 0.5, (0., 10., 0., 10.))
    out = {}
    for f_rew, fn in (('mean', aq_oracle.mean_UCB), ('mes', aq_oracle.mves)):
        planner = mc.cMCTS(31, m, (5., 5., 0.3), 3, pg, fn, f_rew, 2, tree_type='dpw')
        if f_rew == 'mes':                            # host stand-in for the device max-value sampler: counts its draws
            planner.reward_param = lambda t, planner=planner: (np.random.uniform(20, 30, (3, 1)), None, None)
        with contextlib.redirect_stdout(io.StringIO()):
            res = parallel.mcts_root_parallel(planner, 2, seed=123 if f_rew == 'mean' else None)
        out[f_rew + '_action'] = np.array(res[0])
        out[f_rew + '_value'] = np.array([res[2]])
        out[f_rew + '_all'] = np.array([res[4][i] for i in range(len(res[4]))])
        out[f_rew + '_local'] = planner.local_root_stats
        out[f_rew + '_global'] = planner.root_stats
        out[f_rew + '_param'] = np.zeros(1) if planner.tree.param is None else np.asarray(planner.tree.param[0])
    np.savez(os.path.join(tmp, 'mcts%d.npz' % rank), **out)
    dist.destroy_process_group()


def test_root_parallel_tree_search_gloo(tmp_path):
    """parallel.mcts_root_parallel on 2 ranks: the rollout budget is split, each rank searches its own tree with its own
    random stream, the root statistics are summed by one all_reduce and both ranks return the same action."""
    port = _free_port()
    mp.spawn(_mcts_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / 'mcts0.npz'), np.load(tmp_path / 'mcts1.npz')
    for f_rew in ('mean', 'mes'):
        for key in ('_action', '_value', '_all', '_global', '_param'):
            np.testing.assert_array_equal(a[f_rew + key], b[f_rew + key])
        np.testing.assert_allclose(a[f_rew + '_local'] + b[f_rew + '_local'], a[f_rew + '_global'])
        assert a[f_rew + '_local'][:, 0].sum() == 16 and b[f_rew + '_local'][:, 0].sum() == 15      # 31 rollouts split
        assert not np.array_equal(a[f_rew + '_local'], b[f_rew + '_local'])                      # different streams
        g = a[f_rew + '_global']
        means = g[:, 1] / g[:, 0]
        best = sorted(range(len(g)), key=lambda i: (-g[i, 0], -means[i], i))[0]          # most visits, then mean reward
        assert a[f_rew + '_value'][0] == means[best]
        np.testing.assert_allclose(a[f_rew + '_all'], means)
    assert a['mes_param'].shape == (3, 1)


class _FakeBelief(object):
    """Stand-in for the replicated device belief in the sharded max-value sampler: the draws only."""
    variance = 100.0

    def shared_feature_draws(self, nK, nFeatures, rng):
        W = rng.normal(size=(nFeatures, 2))
        b = 2 * np.pi * rng.uniform(size=(nFeatures, 1))
        return W, b, rng.normal(size=(nFeatures, nK))


def _numpy_grid_stage(W, b, Theta, variance, xs, ys):
    gx, gy = np.meshgrid(xs, ys, indexing='xy')
    grid = np.vstack([gx.ravel(), gy.ravel()]).T
    vals = (Theta.T * np.sqrt(2 * variance / W.shape[0])) @ np.cos(W @ grid.T + b)
    return vals.max(axis=1), vals.argmax(axis=1)


def _maxvals_worker(rank, size, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    from informative_path_planning_b200 import parallel
    rng = np.random.default_rng(9)
    xs, ys = rng.uniform(1, 9, 23), rng.uniform(1, 9, 17)           # 17 mesh rows over 2 ranks: 9 + 8
    mx, locs, (W, b, Theta) = parallel.sample_max_vals_shared_sharded(_FakeBelief(), xs, ys, nK=6, nFeatures=40, seed=4,
                                                                      eval_fn=_numpy_grid_stage)
    # ties across ranks go to the smallest global index, like np.argmax over the whole grid
    v = torch.tensor([1.0, 5.0, 2.0]) if rank == 0 else torch.tensor([1.0, 3.0, 7.0])
    i = torch.tensor([40, 11, 12]) if rank == 0 else torch.tensor([7, 50, 3])
    tv, ti = parallel.combine_sharded_argmax(v.to(torch.float64), i.to(torch.int64))
    np.savez(os.path.join(tmp, 'mv%d.npz' % rank), mx=mx, locs=locs, W=W, b=b, Theta=Theta, xs=xs, ys=ys,
             tv=tv.numpy(), ti=ti.numpy())
    dist.destroy_process_group()


def test_sharded_max_value_sampling_gloo(tmp_path):
    """Mesh rows sharded over two ranks, (max, arg-max) pairs gathered and reduced: every rank ends with the maxima and
    locations of the WHOLE mesh, ties resolved like np.argmax."""
    port = _free_port()
    mp.spawn(_maxvals_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'mv0.npz'), np.load(tmp_path / 'mv1.npz')
    for k in ('mx', 'locs', 'W', 'b', 'Theta', 'tv', 'ti'):
        np.testing.assert_array_equal(r0[k], r1[k])
    want_max, want_arg = _numpy_grid_stage(r0['W'], r0['b'], r0['Theta'], 100.0, r0['xs'], r0['ys'])
    np.testing.assert_allclose(r0['mx'].ravel(), want_max, rtol=0, atol=0)
    nx = r0['xs'].shape[0]
    np.testing.assert_array_equal(r0['locs'], np.column_stack([r0['xs'][want_arg % nx], r0['ys'][want_arg // nx]]))
    np.testing.assert_array_equal(r0['tv'], [1.0, 5.0, 7.0])
    np.testing.assert_array_equal(r0['ti'], [7, 11, 3])
