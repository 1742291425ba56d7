#!/bin/bash
# round 2, call n: re-entry check -- GPU suite on the restored tree, profile of the per-step max-value sampling
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2n_pytest.log
timeout 300 python tools/profile_maxvals.py 900 > gpurun_out/r2n_profile_maxvals_900.log 2>&1; echo rc=$?
head -40 gpurun_out/r2n_profile_maxvals_900.log
