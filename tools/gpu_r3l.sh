#!/bin/bash
# round 2, call 3l: mesh form with Phi formed in the A fragments of the GEMM -- parity, A/B timing, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rff or max_vals" > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3l_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r3l_bench_rff.log
IPP_RFF_MESH_FUSED=0 timeout 300 python tools/bench_rff.py 100 2>&1 | tee -a gpurun_out/r3l_bench_rff.log
timeout 300 python tools/bench_rff.py 32 2>&1 | tee -a gpurun_out/r3l_bench_rff.log
python tools/ncu_targets.py rffmesh > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rff_gemm_argmax_kernel -s 2 -c 1 -o gpurun_out/r3l_rff_gemm_mesh -f python tools/ncu_targets.py rffmesh > gpurun_out/r3l_ncu.log 2>&1
