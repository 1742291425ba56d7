"""Information-gain rollouts at N=1000 against the oracle and an extended-precision Schur log-determinant: how far each
route is from the truth (diagnostic for the tolerance of test_info_gain_rollouts_match_sequential_appends)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import informative_path_planning_b200 as pkg
from oracle import aq_oracle as aqo, gpmodel_oracle as gpo
RANGES = (0., 10., 0., 10.)
def world(rng, n, noise):
    X = rng.uniform(0, 10, (n, 2))
    z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, np.sqrt(noise), (n, 1))
    return X, z
aq, L = pkg.aq_library, pkg._lib
for n_obs, sizes in [(1000, [4, 4, 4]), (300, [3, 1, 5, 4])]:
    rng = np.random.default_rng(77 + n_obs)
    X, z = world(rng, n_obs, 0.5)
    d = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=0.5)
    o = gpo.OnlineGPModelOracle(RANGES, 1.0, 100.0, noise=0.5)
    d.add_data(X, z); o.add_data(X, z)
    levels = []
    start = rng.uniform(2, 8, 2)
    for k in sizes:
        pts = start + np.cumsum(rng.normal(0, 0.4, (k, 2)), axis=0)
        start = pts[-1]
        levels.append((pts, rng.normal(0, 10.0, (k, 1)), int(rng.integers(1, 3))))
    got, info = d.rollout_rewards(L.REWARD_INFO, levels)
    fork = copy.copy(d)
    Xa = X.copy()
    for l, (pts, zo, r) in enumerate(levels):
        want_o = float(aqo.info_gain(3, pts, o))
        want_d = float(aq.info_gain(time=3, xvals=pts, robot_model=fork))
        # truth: logdet of the Schur complement of I + vK([Xa; pts]) in long double via Cholesky of the big matrix
        def K(A, B):
            d2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
            return 100.0 * np.exp(-0.5 * d2)
        Z = np.vstack([Xa, pts]).astype(np.longdouble)
        d2 = ((Z[:, None, :] - Z[None, :, :]) ** 2).sum(-1)
        M = np.eye(len(Z), dtype=np.longdouble) + np.longdouble(100.0) * (np.longdouble(100.0) * np.exp(-0.5 * d2))
        # long-double Cholesky (row-oriented, vectorised)
        n = len(Z); Lc = np.zeros_like(M)
        for j in range(n):
            v = M[j:, j] - Lc[j:, :j] @ Lc[j, :j]
            Lc[j:, j] = v / np.sqrt(v[0])
        truth = float(2 * np.pi * np.e * 2.0 * np.sum(np.log(np.diag(Lc)[len(Xa):])))
        print('N=%d level %d: truth %.9f | rollout %+.2e | oracle %+.2e | device per-call %+.2e' % (
            n_obs, l, truth, float(got[l]) - truth, want_o - truth, want_d - truth))
        for _ in range(r):
            fork.add_data(pts, zo); o.add_data(pts, zo); Xa = np.vstack([Xa, pts])
