#!/bin/bash
# round 2, call 4f (2 GPUs): 2-rank equality test incl. sharded max-value sampling
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r4f_pytest_multi.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r4f_pytest_multi.log
