"""Timing of sample_max_vals (aq_library.py:141-279 restated): where does a MES planning step spend its time?"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import informative_path_planning_b200 as pkg
from informative_path_planning_b200 import aq_library as aq
R = (0., 10., 0., 10.)
for N in (100, 900, 4096):
    rng = np.random.default_rng(N)
    X = rng.uniform(0, 10, (N, 2))
    z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (N, 1))
    m = pkg.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
    m.add_data(X, z)
    np.random.seed(1)
    aq.sample_max_vals(m, t=3)
    torch.cuda.synchronize()
    np.random.seed(2)
    t0 = time.perf_counter()
    pr = cProfile.Profile(); pr.enable()
    out = aq.sample_max_vals(m, t=3)
    pr.disable()
    torch.cuda.synchronize()
    print('N=%d sample_max_vals(nK=3, F=200): %.1f ms, maxima %s' % (N, 1e3 * (time.perf_counter() - t0), np.round(out[0].ravel(), 2)))
    if N == 900:
        pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
