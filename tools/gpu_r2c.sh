#!/bin/bash
# round 2, call c: fused rollout kernel + new diagonal-block Cholesky -- parity, then timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2c_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2c_split_fused.log 2>&1; cat gpurun_out/r2c_split_fused.log
IPP_ROLLOUT_TPC=2 timeout 300 python tools/mcts_split.py > gpurun_out/r2c_split_fused_tpc2.log 2>&1; cat gpurun_out/r2c_split_fused_tpc2.log
IPP_ROLLOUT_TPC=4 timeout 300 python tools/mcts_split.py > gpurun_out/r2c_split_fused_tpc4.log 2>&1; cat gpurun_out/r2c_split_fused_tpc4.log
python tools/ncu_targets.py factor > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2c_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
