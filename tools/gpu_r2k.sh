#!/bin/bash
# round 2, call k: rollout v7 (cluster-level partial sums) + new tests; full suite
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2k_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2k_split.log 2>&1; cat gpurun_out/r2k_split.log
