#!/bin/bash
# round 2, call 4a: belief-tree and heuristic-reward search rates with the final kernels
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_mcts_belief.py 2>&1 | tee gpurun_out/r4a_mcts_belief_bench.jsonl
timeout 600 python tools/bench_mcts_naive.py 2>&1 | tail -4 | tee gpurun_out/r4a_mcts_naive_bench.log
