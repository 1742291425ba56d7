#!/bin/bash
# round 2, call 3a: pipelined max-value sampling (device weight draws, one synchronisation per sample), native hotspot search -- full suite, profile, MES step
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r3a_pytest.log
timeout 300 python tools/profile_maxvals.py 900 > gpurun_out/r3a_profile_maxvals_900.log 2>&1; head -30 gpurun_out/r3a_profile_maxvals_900.log
timeout 300 python tools/profile_maxvals.py 4096 > gpurun_out/r3a_profile_maxvals_4096.log 2>&1; head -12 gpurun_out/r3a_profile_maxvals_4096.log
for n in 900 4096; do timeout 300 python tools/mcts_split_n.py $n mes > gpurun_out/r3a_split_${n}_mes.log 2>&1; head -1 gpurun_out/r3a_split_${n}_mes.log; done
