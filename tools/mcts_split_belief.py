"""Where a native belief-tree planning step spends its time (N = 900, UCB): walk / device pass / launch calls, kernel phases."""
import contextlib, ctypes, io, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp, mcts_library as mc, _lib as L
from informative_path_planning_b200.paths_library import Dubins_Path_Generator
R = (0., 10., 0., 10.)
rng = np.random.default_rng(0)
X = rng.uniform(0, 10, (900, 2))
z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (900, 1))
model = gp.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
model.add_data(X, z)
for tree_type in ('dpw', 'belief'):
    best = None
    for rep in range(4):
        pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, R)
        np.random.seed(5); random.seed(6)
        m = mc.cMCTS(250, model, (5.0, 5.0, 0.3), 5, pg, aq.mean_UCB, 'mean', 3, tree_type=tree_type)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
            t0 = time.perf_counter(); m.choose_trajectory(t=3); dt = time.perf_counter() - t0
        w, d, lu = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        L.check(L.lib.ipp_tree_last_timing(ctypes.byref(w), ctypes.byref(d)))
        L.check(L.lib.ipp_tree_last_launch_us(ctypes.byref(lu)))
        if rep > 0 and (best is None or dt < best[0]):
            best = (dt, w.value, d.value, lu.value, m.tree.reward_evals, getattr(m, 'rollout_seconds', dt))
    clk = np.zeros(32, dtype=np.int64)
    L.check(L.lib.ipp_rollout_debug_clocks(clk.ctypes.data_as(ctypes.c_void_p)))
    print('%s: step %.2f ms (rollout loop %.2f ms): walk %.2f + device %.2f (launch calls %.2f) ms, reward evals %d; last kernel phases %s' % (
        tree_type, 1e3 * best[0], 1e3 * best[5], best[1] / 1e3, best[2] / 1e3, best[3] / 1e3, best[4], [int(v) for v in clk[:4]]))
