"""BASELINE config 5: sharded candidate-path scoring -- 65536 (pose, goal) pairs -> Dubins curves generated on the device
(20 samples / path: goal 10.0 ahead in a 40 x 40 world, sample step 0.5) -> UCB / MES rewards against a replicated belief
of N = 4096 observations -> one NCCL all_gather of the per-path rewards.  One process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 tools/bench_config5.py [--precision i8x6]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='fp64')
ap.add_argument('--paths', type=int, default=65536)
args = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from informative_path_planning_b200 import _lib as L, aq_library as aq, gpmodel_library as gp, parallel
from informative_path_planning_b200.paths_library import dubins_paths_device, ragged_points

R = (0., 40., 0., 40.)
N, P = 4096, args.paths
rng = np.random.default_rng(0)
X = rng.uniform(0, 40, (N, 2))
z = 10 * np.sin(0.3 * X[:, :1]) * np.cos(0.2 * X[:, 1:]) + rng.normal(0, 0.7, (N, 1))
model = gp.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
parallel.add_data_replicated(model, X if rank == 0 else None, z if rank == 0 else None)
model.predict_precision = args.precision
poses = np.column_stack([rng.uniform(12, 28, P), rng.uniform(12, 28, P), rng.uniform(-np.pi, np.pi, P)])
ang = poses[:, 2] + rng.uniform(-1.0, 1.0, P)
goals = np.column_stack([poses[:, 0] + 10.0 * np.cos(ang), poses[:, 1] + 10.0 * np.sin(ang), ang])
lo, hi = parallel.shard_bounds(P, rank, world)
poses_d, goals_d = L.to_device(poses[lo:hi]), L.to_device(goals[lo:hi])
maxes = L.to_device(np.linspace(12, 20, 100))
sqb = [float(np.sqrt(aq.ucb_beta(10)))]
width = parallel.shard_bounds(P, 0, world)[1]


def step(kind, **kw):
    samp, counts, _ = dubins_paths_device(poses_d, goals_d, 0.5, 0.5, R, max_samples=24)
    pts, offs, kept = ragged_points(samp, counts)
    rew, _ = model.reward_paths_device(kind, pts, offs, **kw)
    full = torch.full((width,), float('nan'), dtype=torch.float64, device=rew.device)    # dropped paths stay NaN
    full[kept] = rew
    if world > 1:
        out = torch.empty(world * width, dtype=torch.float64, device=rew.device)
        dist.all_gather_into_tensor(out, full)
        full = out
    return full, int(pts.shape[0]), int(kept.numel())


res = {}
for name, kind, kw in (('ucb', L.REWARD_UCB, dict(params=sqb)), ('mes_S100', L.REWARD_MES, dict(maxes_d=maxes))):
    for _ in range(2):
        step(kind, **kw)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        full, n_pts, n_kept = step(kind, **kw)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3, float(n_pts), float(n_kept)], dtype=torch.float64, device='cuda')
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    else:
        t = t.clone()
    res[name] = {'ms_per_pass_max_over_ranks': float(tmax[0]), 'points': int(t[1]), 'kept_paths': int(t[2]),
                 'reward_evals_per_s': float(t[2]) / (float(tmax[0]) / 1e3), 'finite_rewards': int(torch.isfinite(full).sum())}
if rank == 0:
    print(json.dumps({'config': 'BASELINE config 5: %d candidate Dubins paths (device-generated, ~20 samples each), N=4096, '
                                'paths sharded over %d GPU(s), rewards all-gathered' % (P, world), 'n_gpus': world,
                      'precision': args.precision, **res}))
if world > 1:
    dist.destroy_process_group()
