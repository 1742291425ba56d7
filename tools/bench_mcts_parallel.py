"""Root-parallel tree search across the GPUs of one box (BASELINE config 2 sharding):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/bench_mcts_parallel.py [n_obs] [f_rew]
Every rank holds the replicated belief (N observations) and runs 250 rollouts (dpw tree, horizon 5, Dubins fan) of ONE
planning step on its own GPU; the root statistics are summed with one all_reduce (parallel.mcts_root_parallel).
Weak scaling: 250 rollouts per rank.  Wall-clock per planning step (host tree + device calls), max over ranks."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

R = (0., 10., 0., 10.)


def main():
    n_obs = int(sys.argv[1]) if len(sys.argv) > 1 else 900
    f_rew = sys.argv[2] if len(sys.argv) > 2 else 'mean'
    rank = int(os.environ.get('RANK', 0)); size = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if size > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp, mcts_library as mc, parallel
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 10, (n_obs, 2))
    z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (n_obs, 1))
    model = gp.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
    parallel.add_data_replicated(model, X if rank == 0 else None, z if rank == 0 else None)
    pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, R)
    fn = {'mean': aq.mean_UCB, 'mes': aq.mves}[f_rew]
    per_rank = 250
    best = None
    for rep in range(4):                                   # first repetition = warm-up
        planner = mc.cMCTS(per_rank * size, model, (5.0, 5.0, 0.3), 5, pg, fn, f_rew, 3, tree_type='dpw')
        if size > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
            res = parallel.mcts_root_parallel(planner, 3, seed=100 + rep)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
        if size > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep > 0:
            best = float(dt.item()) if best is None else min(best, float(dt.item()))
    if rank == 0:
        visits = planner.root_stats[:, 0] if size > 1 else np.array([c.nqueries for c in planner.tree.root.children])
        print(json.dumps({'metric': 'mcts_rollouts_per_s', 'n_gpus': size, 'n_obs': n_obs, 'f_rew': f_rew,
                          'rollouts_per_step': per_rank * size, 'seconds_per_step': best,
                          'value': per_rank * size / best, 'scaling': 'weak', 'root_visits': visits.tolist(),
                          'timing': 'wall clock of one planning step, max over ranks, best of 3'}))
    if size > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
