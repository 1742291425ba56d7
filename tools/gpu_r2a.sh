#!/bin/bash
# round 2, call a: baseline of the round (tests, bench) + the ncu captures VERDICT r1 asked for
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cap() {  # name target kernel-regex skip
  python tools/ncu_targets.py $2 > gpurun_out/r2a_plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o gpurun_out/r2a_$1 -f python tools/ncu_targets.py $2 > gpurun_out/r2a_ncu_$1.log 2>&1
  echo "cap $1 rc=$?"
}
cap rbf_queries rbf '^rbf_queries_kernel' 1
cap rbf_queries_i8 rbf_i8 rbf_queries_i8_kernel 1
cap rff_gemm_argmax rff rff_gemm_argmax_kernel 1
cap append_update append append_update_kernel 1
cap info_gain_schur infogain info_gain_schur_kernel 1
cap rollout_condition rollout rollout_condition_kernel 20
cap potrf_diag factor potrf_diag_kernel 40
python tools/ncu_targets.py rollout > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_rollout900.csv python tools/ncu_targets.py rollout > /dev/null 2>&1
python tools/ncu_targets.py factor > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2a_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
ls -la gpurun_out
