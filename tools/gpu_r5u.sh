#!/bin/bash
# round 2, call 5u: general-grid feature pass with the magnitude test hoisted per feature
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rff or max_vals or sample_max or mesh or naive" > gpurun_out/r5u_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r5u_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r5u_bench_rff.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rff_phi_kernel -c 4 --csv --log-file gpurun_out/r5u_launches_phi.csv python tools/ncu_targets.py rff 1 > /dev/null 2>&1; grep rff_phi gpurun_out/r5u_launches_phi.csv | awk -F, '{print $NF}' | tr '\n' ' '; echo
