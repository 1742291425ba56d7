#!/bin/bash
# round 2, call 5m: mesh-fused grid-stage GEMM with the phase-shifted lean loader
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rff or max_vals or sample_max or mesh" > gpurun_out/r5m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5m_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r5m_bench_rff.log
timeout 300 python tools/bench_rff.py 32 2>&1 | tee -a gpurun_out/r5m_bench_rff.log
