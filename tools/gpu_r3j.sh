#!/bin/bash
# round 2, call 3j: two-level blocked Cholesky (rank-64 updates inside a 256-column panel, rank-256 trailing updates) -- parity, refresh times, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_precision.py -m gpu -q -x -k "pdinv or factor or golden or append or max_vals or potrf or precision or oracle" > gpurun_out/r3j_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3j_pytest.log
python tools/ncu_targets.py factor > gpurun_out/r3j_factor_clocks.log 2>&1; cat gpurun_out/r3j_factor_clocks.log
python tools/bench_refresh.py 2>&1 | tee gpurun_out/r3j_refresh_ms.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r3j_launches_factor_4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3j_launches_factor_4096.csv > gpurun_out/r3j_factor_4096_launch_summary.txt; head -4 gpurun_out/r3j_factor_4096_launch_summary.txt
