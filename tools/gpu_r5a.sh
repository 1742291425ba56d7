#!/bin/bash
# round 2, call 5a: one-wave rbf_queries_kernel -- parity tests that run through it, launch list, ncu full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or rbf or info or golden" > gpurun_out/r5a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5a_pytest.log
python tools/ncu_targets.py rbf 5 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r5a_launches_rbf.csv python tools/ncu_targets.py rbf 5 > gpurun_out/r5a_ncu_list.log 2>&1
grep rbf_queries gpurun_out/r5a_launches_rbf.csv | tail -5
ncu --set full --clock-control none --import-source on -k regex:rbf_queries_kernel -s 2 -c 1 -o gpurun_out/r5a_rbf_queries -f python tools/ncu_targets.py rbf > gpurun_out/r5a_ncu_full.log 2>&1
ls -la gpurun_out/r5a_rbf_queries.ncu-rep
