#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "theta_draw" > gpurun_out/r3r_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r3r_pytest.log
