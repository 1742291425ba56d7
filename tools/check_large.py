"""Largest BASELINE size (N = 8192 observations): every precision mode against the CPU oracle on a slice of queries."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import informative_path_planning_b200 as pkg
from oracle.gpmodel_oracle import OnlineGPModelOracle
R = (0., 10., 0., 10.)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(N)
X = rng.uniform(0, 10, (N, 2))
z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, np.sqrt(0.5), (N, 1))
d = pkg.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
t0 = time.perf_counter(); d.add_data(X, z); t_dev = time.perf_counter() - t0
o = OnlineGPModelOracle(R, 1.0, 100.0, noise=0.5)
t0 = time.perf_counter(); o.add_data(X, z); t_cpu = time.perf_counter() - t0
q = rng.uniform(0, 10, (4096, 2))
mu_o, var_o = o.predict_value(q[:512])
out = {'N': N, 'factor_device_s': t_dev, 'factor_cpu_s': t_cpu}
res = {}
for prec in ('fp64', 'i8x7', 'i8x6', 'i8x5'):
    d.predict_precision = prec
    mu, var = d.predict_value(q)
    res[prec] = var
    out[prec] = {'mean_rel_err_max': float(np.max(np.abs(mu[:512] - mu_o) / (np.abs(mu_o) + 1e-6))),
                 'var_rel_err_max': float(np.max(np.abs(var[:512] - var_o) / var_o))}
# extended-precision arbiter on a few queries: Cholesky solve in float64 + iterative refinement with long-double
# residuals, so that q = k^T (K + s I)^-1 k is good to ~1e-15 and the modes can be ranked against the truth
import scipy.linalg as sl
nq = 24
def Kf(A, B):
    return 100.0 * np.exp(-0.5 * ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
Ky = Kf(X, X) + (0.5 + 1e-8) * np.eye(N)
c = sl.cho_factor(Ky, lower=True)
Ky_l = Ky.astype(np.longdouble)
kq = Kf(X, q[:nq])                                   # N x nq
x = sl.cho_solve(c, kq).astype(np.longdouble)
for _ in range(3):
    r = kq.astype(np.longdouble) - Ky_l @ x
    x = x + sl.cho_solve(c, r.astype(np.float64)).astype(np.longdouble)
truth = (np.longdouble(100.0) - np.einsum('ij,ij->j', kq.astype(np.longdouble), x) + np.longdouble(0.5)).astype(np.float64)
out['vs_extended_precision_truth_%d_queries' % nq] = {
    p: float(np.max(np.abs(res[p][:nq, 0] - truth) / truth)) for p in res}
out['vs_extended_precision_truth_%d_queries' % nq]['oracle_numpy'] = float(np.max(np.abs(var_o[:nq, 0] - truth) / truth))
print(json.dumps(out, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/check_large_N%d.json' % N, 'w'), indent=1)
