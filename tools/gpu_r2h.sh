#!/bin/bash
# round 2, call h: rollout v6 (fixed-order partials, trimmed setup / images / conditioning), potrf clocks
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2h_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2h_split.log 2>&1; cat gpurun_out/r2h_split.log
python tools/ncu_targets.py factor > gpurun_out/r2h_factor_clocks.log 2>&1; cat gpurun_out/r2h_factor_clocks.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2h_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2h_launches_factor4096.csv > gpurun_out/r2h_factor_summary.txt; head -4 gpurun_out/r2h_factor_summary.txt
