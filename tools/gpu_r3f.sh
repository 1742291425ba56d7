#!/bin/bash
# round 2, call 3f: grid stage of max-value sampling with the features generated inside the GEMM -- parity, A/B timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rff or max_vals or mves or mes" > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3f_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r3f_bench_rff.log
IPP_RFF_FUSED=0 timeout 300 python tools/bench_rff.py 100 2>&1 | tee -a gpurun_out/r3f_bench_rff.log
timeout 300 python tools/bench_rff.py 128 2>&1 | tee -a gpurun_out/r3f_bench_rff.log
timeout 300 python tools/bench_rff.py 32 2>&1 | tee -a gpurun_out/r3f_bench_rff.log
python tools/ncu_targets.py rff > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rff_gemm_argmax_kernel -s 40 -c 1 -o gpurun_out/r3f_rff_gemm_fused -f python tools/ncu_targets.py rff > gpurun_out/r3f_ncu_rff.log 2>&1
