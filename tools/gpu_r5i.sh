#!/bin/bash
# round 2, call 5i: rollout kernel with the exp coefficients as constant-bank operands (variant build) against the current one
set -x
mkdir -p gpurun_out
V=$PWD/informative_path_planning_b200/csrc/libipp_b200_expcb.so
for rep in 1 2; do
for n in 900 4096; do
  timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r5i_split_${n}_base_$rep.log 2>&1; head -3 gpurun_out/r5i_split_${n}_base_$rep.log
  IPP_B200_LIB=$V timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r5i_split_${n}_expcb_$rep.log 2>&1; head -3 gpurun_out/r5i_split_${n}_expcb_$rep.log
done
done
IPP_B200_LIB=$V timeout 900 python -m pytest tests -m gpu -q -x -k "tree or rollout or mission or mcts" > gpurun_out/r5i_pytest_expcb.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5i_pytest_expcb.log
