#!/bin/bash
# round 2, call 3x: 64 x 64-tile GEMM for launches that do not fill the GPU -- parity (everything that factors), refresh times, A/B
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_precision.py tests/test_gpu_mission.py -m gpu -q -x > gpurun_out/r3x_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3x_pytest.log
python tools/bench_refresh.py 2>&1 | tee gpurun_out/r3x_refresh_ms.txt
IPP_GEMM_SMALL=0 python tools/bench_refresh.py 2>&1 | sed 's/^/big tiles only: /' | tee -a gpurun_out/r3x_refresh_ms.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r3x_launches_factor_4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3x_launches_factor_4096.csv > gpurun_out/r3x_factor_4096_launch_summary.txt; head -5 gpurun_out/r3x_factor_4096_launch_summary.txt
timeout 300 python tools/profile_maxvals.py 900 2>&1 | head -8 | tail -3
