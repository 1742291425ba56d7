"""Belief refresh (K + noise I -> Cholesky -> L^-1 -> W^-1 -> alpha) at N = 1000 / 4096 / 8192, CUDA events, mean of 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from informative_path_planning_b200 import gpmodel_library as gp
rng = np.random.default_rng(0)
for N in (200, 1000, 4096, 8192):
    X = rng.uniform(0, 10, (N, 2)); z = rng.normal(0, 10, (N, 1))
    m = gp.OnlineGPModel((0., 10., 0., 10.), 1.0, 100.0, noise=0.5); m.add_data(X, z)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m._factor(want_linv=False)
    e1.record(); torch.cuda.synchronize()
    print('refresh N=%d: %.2f ms' % (N, e0.elapsed_time(e1) / 3))
