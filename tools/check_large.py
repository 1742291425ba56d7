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
for prec in ('fp64', 'i8x6', 'i8x5', 'tf32x3'):
    d.predict_precision = prec
    mu, var = d.predict_value(q)
    out[prec] = {'mean_rel_err_max': float(np.max(np.abs(mu[:512] - mu_o) / (np.abs(mu_o) + 1e-6))),
                 'var_rel_err_max': float(np.max(np.abs(var[:512] - var_o) / var_o))}
print(json.dumps(out, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/check_large_N%d.json' % N, 'w'), indent=1)
