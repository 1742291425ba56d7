#!/bin/bash
# round 2, call 5h: cos with both kernel polynomials evaluated (general-grid feature pass); rff / max-value tests; ncu of the feature pass
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rff or max_vals or sample_max or mes or naive" > gpurun_out/r5h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5h_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r5h_bench_rff.log
python tools/ncu_targets.py rff > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rff_phi_kernel -s 2 -c 1 -o gpurun_out/r5h_rff_phi -f python tools/ncu_targets.py rff > gpurun_out/r5h_ncu.log 2>&1
ls -la gpurun_out/r5h_rff_phi.ncu-rep
