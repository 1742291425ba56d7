#!/bin/bash
set -x
mkdir -p gpurun_out
for n in 900; do timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r3v_split_$n.log 2>&1; tail -5 gpurun_out/r3v_split_$n.log; done
timeout 300 python tools/mcts_split_n.py 900 mes > gpurun_out/r3v_split_900_mes.log 2>&1; tail -5 gpurun_out/r3v_split_900_mes.log
