#!/bin/bash
# round 2, call u: grid stage of max-value sampling -- 8 NB-wide tile (S=100 -> 104 columns), fast cos in the feature pass
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rff or max_vals or mves or mes" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2u_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r2u_bench_rff.log
timeout 300 python tools/bench_rff.py 128 2>&1 | tee -a gpurun_out/r2u_bench_rff.log
python tools/ncu_targets.py rff > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches_rff.csv python tools/ncu_targets.py rff > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2u_launches_rff.csv | head
