#!/bin/bash
# round 2, call 5c: rbf_queries_kernel with operand prefetch -- variants, ncu full
set -x
mkdir -p gpurun_out
for v in "1 3" "0 3" "1 2" "0 2"; do set -- $v
  IPP_RBFQ_WAVES=$1 IPP_RBFQ_OCC=$2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rbf_queries_kernel -c 6 --csv --log-file gpurun_out/r5c_launches_rbf_w$1_o$2.csv python tools/ncu_targets.py rbf 6 > /dev/null 2>&1
  echo "waves=$1 occ=$2"; grep rbf_queries gpurun_out/r5c_launches_rbf_w$1_o$2.csv | awk -F, '{print $NF}' | tr '\n' ' '; echo
done
ncu --set full --clock-control none --import-source on -k regex:rbf_queries_kernel -s 2 -c 1 -o gpurun_out/r5c_rbf_queries -f python tools/ncu_targets.py rbf > gpurun_out/r5c_ncu_full.log 2>&1
ls -la gpurun_out/r5c_rbf_queries.ncu-rep
