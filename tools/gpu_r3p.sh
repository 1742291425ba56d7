#!/bin/bash
# round 2, call 3p: one-launch rollout, cluster mates push their partials (one cluster barrier instead of two) -- parity, clocks
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout or mcts or tree" > gpurun_out/r3p_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3p_pytest.log
for n in 100 900; do timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r3p_split_$n.log 2>&1; tail -4 gpurun_out/r3p_split_$n.log; done
timeout 300 python tools/mcts_split_n.py 900 mes > gpurun_out/r3p_split_900_mes.log 2>&1; tail -4 gpurun_out/r3p_split_900_mes.log
