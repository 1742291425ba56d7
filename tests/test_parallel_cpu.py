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
