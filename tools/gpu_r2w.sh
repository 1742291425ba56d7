#!/bin/bash
# round 2, call w: information-gain rollouts against an extended-precision truth, old and new diagonal-block kernels
set -x
mkdir -p gpurun_out
timeout 600 python tools/diag_ig.py > gpurun_out/r2w_diag_ig_new.log 2>&1; cat gpurun_out/r2w_diag_ig_new.log | tail -8
IPP_B200_LIB=$PWD/informative_path_planning_b200/csrc/libipp_b200_oldpotrf.so timeout 600 python tools/diag_ig.py > gpurun_out/r2w_diag_ig_old.log 2>&1; cat gpurun_out/r2w_diag_ig_old.log | tail -8
