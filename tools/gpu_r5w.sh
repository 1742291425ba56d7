#!/bin/bash
# round 2, call 5w: ncu full capture of the mesh-fused grid-stage GEMM after the loader changes
set -x
mkdir -p gpurun_out
python tools/ncu_targets.py rffmesh 1 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rff_gemm_argmax_kernel -s 1 -c 1 -o gpurun_out/r5w_rff_gemm_mesh -f python tools/ncu_targets.py rffmesh 2 > gpurun_out/r5w_ncu.log 2>&1
ls -la gpurun_out/r5w_rff_gemm_mesh.ncu-rep
