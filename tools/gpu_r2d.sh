#!/bin/bash
# round 2, call d: full GPU suite (fused rollout, info-gain rollouts, new potrf), phase clocks, precision-mode table
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2d_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2d_split.log 2>&1; cat gpurun_out/r2d_split.log
timeout 600 python tools/check_precision_modes.py > gpurun_out/r2d_precision.log 2>&1; cat gpurun_out/r2d_precision.log
