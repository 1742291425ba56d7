#!/bin/bash
# round 2, call 3u: one-launch rollout, small-block conditioning by the first warp only -- parity, clocks
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout or mcts or tree" > gpurun_out/r3u_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3u_pytest.log
for n in 100 900; do timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r3u_split_$n.log 2>&1; tail -4 gpurun_out/r3u_split_$n.log; done
timeout 300 python tools/mcts_split_n.py 900 mes > gpurun_out/r3u_split_900_mes.log 2>&1; tail -4 gpurun_out/r3u_split_900_mes.log
