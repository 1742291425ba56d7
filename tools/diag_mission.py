import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, scipy.linalg as sl
import informative_path_planning_b200 as pkg
from informative_path_planning_b200 import mcts_library as mc
from informative_path_planning_b200.paths_library import Dubins_Path_Generator
from oracle import aq_oracle
from oracle.gpmodel_oracle import OnlineGPModelOracle
from test_gpu_mission import field, RANGES
noise = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-4
d = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=noise); o = OnlineGPModelOracle(RANGES, 1.0, 100.0, noise=noise)
rng = np.random.default_rng(3)
x0 = rng.uniform(1, 9, (5, 2)); z0 = field(x0) + rng.normal(0, np.sqrt(noise), (5, 1))
d.add_data(x0, z0); o.add_data(x0, z0)
pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, RANGES)
pose = (5.0, 5.0, 0.0)
def K(A, B):
    dd = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1); return 100 * np.exp(-0.5 * dd)
for t in range(40):
    paths, true_paths = pg.get_path_set(pose)
    np.random.seed(1000 + t)
    best = mc.choose_myopic(d, pg, pose, 'mean', t)
    vals = {k: float(aq_oracle.mean_UCB(t, paths[k], o)) for k in paths}
    X, z = d.xvals, d.zvals
    c = sl.cho_factor(K(X, X) + (noise + 1e-8) * np.eye(len(X)), lower=True)
    truth = {}
    for k in paths:
        q = np.array(paths[k])[:, :2]; kx = K(X, q)
        mu = kx.T @ sl.cho_solve(c, z); var = 100 - np.einsum('ij,ij->j', kx, sl.cho_solve(c, kx)) + noise
        truth[k] = float(mu.sum() + np.sqrt(aq_oracle.ucb_beta(t)) * np.abs(var).sum())
    ks = sorted(paths)
    dv = np.array([best[4][k] for k in ks]); ov = np.array([vals[k] for k in ks]); tv = np.array([truth[k] for k in ks])
    kd = ks[int(np.argmax(dv))]; ko = ks[int(np.argmax(ov))]; kt = ks[int(np.argmax(tv))]
    print('t=%2d N=%3d dev-truth %.1e  ora-truth %.1e  argmax dev/ora/truth %d %d %d' % (t, len(X), np.abs(dv - tv).max(), np.abs(ov - tv).max(), kd, ko, kt))
    xobs = np.array(paths[kd])[:, :2]; zobs = field(xobs) + rng.normal(0, np.sqrt(noise), (xobs.shape[0], 1))
    d.add_data(xobs, zobs); o.add_data(xobs, zobs)
    pose = true_paths[kd][-1]
