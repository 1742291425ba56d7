"""i8x6 predict with the A slices staged in tensor memory (IPP_I8_TS=1) or read from shared memory: variance against the FP64 mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from informative_path_planning_b200 import gpmodel_library as gp, _lib as L
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
X = rng.uniform(0, 10, (N, 2)); z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (N, 1))
m = gp.OnlineGPModel((0., 10., 0., 10.), 1.0, 100.0, noise=0.5); m.add_data(X, z)
q = L.to_device(rng.uniform(0, 10, (40000, 2)))
mu64, v64 = m.predict_device(q, precision='fp64')
mu8, v8 = m.predict_device(q, precision='i8x6')
torch.cuda.synchronize()
rel = ((v8 - v64).abs() / v64.abs()).max().item()
print('IPP_I8_TS=%s N=%d: i8x6 variance max rel err vs fp64 %.3e, mean max abs diff %.3e' % (os.environ.get('IPP_I8_TS', '0'), N, rel, (mu8 - mu64).abs().max().item()))
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.predict_device(q, precision='i8x6'); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print('  40000 queries: %.3f ms' % min(ts))
