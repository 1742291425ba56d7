"""Every predict mode against an extended-precision solve (Cholesky + iterative refinement with long-double residuals),
N in {1024, 4096}, noise in {0.5, 1e-4}: the measurement behind tests/test_gpu_precision.py's asserted bounds."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from test_gpu_precision import extended_precision_variance, make_case
import informative_path_planning_b200 as pkg
out = []
for N in (1024, 4096):
    for noise in (0.5, 1e-4):
        X, z, q = make_case(N, noise)
        truth = extended_precision_variance(X, q, noise)
        d = pkg.OnlineGPModel((0., 10., 0., 10.), 1.0, 100.0, noise=noise)
        d.add_data(X, z)
        row = {'N': N, 'noise': noise, 'truth_min': float(truth.min()), 'truth_max': float(truth.max())}
        big = np.vstack([q, np.random.default_rng(1).uniform(0, 10, (4096, 2))])       # > 64 queries: the tensor modes engage
        for prec in ('fp64', 'i8x7', 'i8x6', 'i8x5'):
            d.predict_precision = prec
            var = d.predict_value(big)[1][:q.shape[0], 0]
            row[prec] = {'rel_max': float(np.max(np.abs(var - truth) / truth)), 'abs_over_prior_max': float(np.max(np.abs(var - truth)) / 100.0)}
        out.append(row)
        print(json.dumps(row))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/precision_modes.json', 'w'), indent=1)
