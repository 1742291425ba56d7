"""cProfile of aq_library.sample_max_vals (the per-step max-value sampling of an MES planning step)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 900
rng = np.random.default_rng(0)
X = rng.uniform(0, 10, (n, 2))
z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (n, 1))
m = gp.OnlineGPModel((0., 10., 0., 10.), 1.0, 100.0, noise=0.5)
m.add_data(X, z)
np.random.seed(0)
aq.sample_max_vals(m, 3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    aq.sample_max_vals(m, 3)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
