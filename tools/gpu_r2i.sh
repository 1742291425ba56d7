#!/bin/bash
# round 2, call i: batched triangular-inverse GEMMs, potrf with uniform block skipping; full suite; factor launch list
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2i_pytest.log
python tools/ncu_targets.py factor > gpurun_out/r2i_factor_clocks.log 2>&1; cat gpurun_out/r2i_factor_clocks.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2i_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2i_launches_factor4096.csv > gpurun_out/r2i_factor_summary.txt; head -4 gpurun_out/r2i_factor_summary.txt
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from informative_path_planning_b200 import gpmodel_library as gp
rng = np.random.default_rng(0)
for N in (1000, 4096, 8192):
    X = rng.uniform(0, 10, (N, 2)); z = rng.normal(0, 10, (N, 1))
    m = gp.OnlineGPModel((0., 10., 0., 10.), 1.0, 100.0, noise=0.5); m.add_data(X, z)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m._factor(want_linv=False)
    e1.record(); torch.cuda.synchronize()
    print('refresh N=%d: %.2f ms' % (N, e0.elapsed_time(e1) / 3))
PY
