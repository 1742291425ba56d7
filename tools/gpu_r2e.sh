#!/bin/bash
# round 2, call e: fused rollout v3 (tagged records, single-warp conditioning), full GPU suite, phase clocks, factor launches
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2e_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2e_split.log 2>&1; cat gpurun_out/r2e_split.log
python tools/ncu_targets.py factor > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2e_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2e_launches_factor4096.csv | head -4
