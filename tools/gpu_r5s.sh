#!/bin/bash
# round 2, call 5s: ncu full capture of predict_var_kernel after the loader change
set -x
mkdir -p gpurun_out
python tools/ncu_targets.py rbf > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:predict_var_kernel -s 1 -c 1 -o gpurun_out/r5s_predict_var -f python tools/ncu_targets.py rbf > gpurun_out/r5s_ncu_full.log 2>&1
ls -la gpurun_out/r5s_predict_var.ncu-rep
