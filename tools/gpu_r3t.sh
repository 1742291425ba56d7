#!/bin/bash
# round 2, call 3t: final-state evidence -- full suite, smoke, bench N=1, ncu launch list of the bench, rollout launch lists
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3t_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3t_bench.json 2> gpurun_out/r3t_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3t_bench.err
python bench.py --steps 2 --warmup 1 --no-extras --queries 131072 > gpurun_out/r3t_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3t_launches_bench_q131072.csv python bench.py --steps 2 --warmup 1 --no-extras --queries 131072 > gpurun_out/r3t_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r3t_launches_bench_q131072.csv > gpurun_out/r3t_bench_launch_summary.txt; head -8 gpurun_out/r3t_bench_launch_summary.txt
python tools/ncu_targets.py rollout > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r3t_launches_mcts_native_900.csv python tools/ncu_targets.py rollout > /dev/null 2>&1
python tools/ncu_targets.py rollout4096 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3t_launches_mcts_native_4096.csv python tools/ncu_targets.py rollout4096 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3t_launches_mcts_native_900.csv | grep -i rollout
python tools/launch_summary.py gpurun_out/r3t_launches_mcts_native_4096.csv | grep -i rollout
