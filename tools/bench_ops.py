"""Per-call latency of the drop-in API at the reference's real problem sizes (k=4 points, N~900)."""
import copy, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp, _lib as L
R = (0., 10., 0., 10.)
rng = np.random.default_rng(0)
out = {}
for N in (100, 900, 4096):
    X = rng.uniform(0, 10, (N, 2)); z = rng.normal(0, 10, (N, 1))
    m = gp.OnlineGPModel(R, 1.0, 100.0, noise=0.5); m.add_data(X, z)
    q = rng.uniform(1, 9, (4, 2)); zq = rng.normal(0, 10, (4, 1))
    for _ in range(20): aq.mean_UCB(time=1, xvals=q, robot_model=m)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300): aq.mean_UCB(time=1, xvals=q, robot_model=m)
    t_ucb = (time.perf_counter() - t0) / 300
    for _ in range(10): f = copy.copy(m); f.add_data(q, zq)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): f = copy.copy(m); f.add_data(q, zq)
    torch.cuda.synchronize(); t_app = (time.perf_counter() - t0) / 100
    # device-only time of the same two ops (CUDA events, no host sync inside)
    xq = L.to_device(q); offs = torch.tensor([0, 4], dtype=torch.int64, device='cuda')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): m.reward_paths_device(L.REWARD_UCB, xq, offs, params=[3.0])
    e1.record(); torch.cuda.synchronize(); d_ucb = e0.elapsed_time(e1) / 100 * 1e-3
    out[N] = {'mean_UCB_call_us': t_ucb * 1e6, 'add_data_call_us': t_app * 1e6, 'reward_device_stream_us': d_ucb * 1e6}
    print(N, out[N])
json.dump(out, open('gpurun_out/ops_latency.json', 'w'), indent=1)
