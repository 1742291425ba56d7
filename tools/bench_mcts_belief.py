"""Belief-tree search (tree_type='belief'): rollouts/s of one planning step, native vs Python tree (N = 900, UCB and MES)."""
import contextlib, io, json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp, mcts_library as mc
from informative_path_planning_b200.paths_library import Dubins_Path_Generator
R = (0., 10., 0., 10.)
rng = np.random.default_rng(0)
X = rng.uniform(0, 10, (900, 2))
z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (900, 1))
model = gp.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
model.add_data(X, z)
for f_rew, fn in (('mean', aq.mean_UCB), ('mes', aq.mves)):
    rows = {}
    for native in (True, False):
        best = None
        for rep in range(4):
            pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, R)
            np.random.seed(5); random.seed(6)
            m = mc.cMCTS(250, model, (5.0, 5.0, 0.3), 5, pg, fn, f_rew, 3, tree_type='belief')
            m.NATIVE_SEARCH = native
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
                t0 = time.perf_counter(); m.choose_trajectory(t=3); dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
        rows[native] = (best, [c.nqueries for c in m.tree.root.children], m.native_search_used)
    print(json.dumps({'tree': 'belief', 'n_obs': 900, 'f_rew': f_rew, 'native_rollouts_per_s': 250 / rows[True][0],
                      'python_tree_rollouts_per_s': 250 / rows[False][0], 'same_root_visits': rows[True][1] == rows[False][1],
                      'native_used': rows[True][2]}))
