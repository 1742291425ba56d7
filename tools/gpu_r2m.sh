#!/bin/bash
# round 2, call m: final-state evidence -- full suite, bench N=1, ncu launch list of the bench, ncu full of the rollout kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2m_bench.err
python bench.py --steps 2 --warmup 1 --no-extras --queries 131072 > gpurun_out/r2m_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_bench_q131072.csv python bench.py --steps 2 --warmup 1 --no-extras --queries 131072 > gpurun_out/r2m_ncu_bench.log 2>&1
python tools/ncu_targets.py rollout > gpurun_out/r2m_plain_rollout.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout_fused_kernel -s 20 -c 1 -o gpurun_out/r2m_rollout_fused -f python tools/ncu_targets.py rollout > gpurun_out/r2m_ncu_rollout.log 2>&1
python tools/ncu_targets.py rollout4096 > gpurun_out/r2m_plain_rollout4096.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout_fused_kernel -s 20 -c 1 -o gpurun_out/r2m_rollout_fused_4096 -f python tools/ncu_targets.py rollout4096 > gpurun_out/r2m_ncu_rollout4096.log 2>&1
python tools/ncu_targets.py rollout > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2m_launches_rollout900.csv python tools/ncu_targets.py rollout > /dev/null 2>&1
timeout 300 python tools/mcts_split.py > gpurun_out/r2m_split.log 2>&1; grep "N=" gpurun_out/r2m_split.log
ls -la gpurun_out | grep r2m
