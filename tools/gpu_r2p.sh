#!/bin/bash
# round 2, call p: streaming rollout kernel -- where the issue clocks go
set -x
mkdir -p gpurun_out
for n in 4096; do timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r2p_split_$n.log 2>&1; tail -4 gpurun_out/r2p_split_$n.log; done
