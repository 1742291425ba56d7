#!/bin/bash
# round 2, call 3k: ncu --set full captures of the kernels changed this session
set -x
mkdir -p gpurun_out
cap() {  # name, kernel regex, target, skip
  python tools/ncu_targets.py $3 > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o gpurun_out/r3k_$1 -f python tools/ncu_targets.py $3 > gpurun_out/r3k_ncu_$1.log 2>&1
  ls -la gpurun_out/r3k_$1.ncu-rep
}
cap potrf_diag potrf_diag_kernel factor 70
cap append_pq append_pq_kernel append 1
cap rff_gemm13 rff_gemm_argmax_kernel rffmesh 30
cap rff_phi_mesh rff_phi_mesh_kernel rffmesh 30
cap rollout_fused_900 rollout_fused_kernel rollout 20
