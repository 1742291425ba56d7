#!/bin/bash
# round 2, call 5d: rbf_queries_kernel staged in shared memory vs the global-load forms; ncu full; parity tests
set -x
mkdir -p gpurun_out
for v in 2 1 0; do
  IPP_RBFQ_MODE=$v ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rbf_queries_kernel -c 6 --csv --log-file gpurun_out/r5d_launches_rbf_mode$v.csv python tools/ncu_targets.py rbf 6 > /dev/null 2>&1
  echo "mode=$v"; grep rbf_queries gpurun_out/r5d_launches_rbf_mode$v.csv | awk -F, '{print $NF}' | tr '\n' ' '; echo
done
ncu --set full --clock-control none --import-source on -k regex:rbf_queries_kernel -s 2 -c 1 -o gpurun_out/r5d_rbf_queries -f python tools/ncu_targets.py rbf > gpurun_out/r5d_ncu_full.log 2>&1
ls -la gpurun_out/r5d_rbf_queries.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or rbf or info or golden" > gpurun_out/r5d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5d_pytest.log
