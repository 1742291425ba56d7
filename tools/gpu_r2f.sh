#!/bin/bash
# round 2, call f: fused rollout v4 (fixed-point accumulation, fast rsqrt), potrf with fast rsqrt + staged loads
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2f_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2f_split.log 2>&1; cat gpurun_out/r2f_split.log
python tools/ncu_targets.py factor > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2f_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2f_launches_factor4096.csv > gpurun_out/r2f_factor_summary.txt; head -4 gpurun_out/r2f_factor_summary.txt
