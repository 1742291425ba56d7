#!/bin/bash
# round 2, call t: streaming rollout kernel (producer warp, tensor-map panels, two accumulator chains) -- parity, rates, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout or mcts or tree" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2t_pytest.log
for n in 900 2048 4096; do timeout 300 python tools/mcts_split_n.py $n > gpurun_out/r2t_split_$n.log 2>&1; tail -4 gpurun_out/r2t_split_$n.log; done
timeout 300 python tools/mcts_split_n.py 4096 mes > gpurun_out/r2t_split_4096_mes.log 2>&1; tail -4 gpurun_out/r2t_split_4096_mes.log
python tools/ncu_targets.py rollout4096 > gpurun_out/r2t_plain_rollout4096.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout_fused_kernel -s 20 -c 1 -o gpurun_out/r2t_rollout_fused_4096 -f python tools/ncu_targets.py rollout4096 > gpurun_out/r2t_ncu_rollout4096.log 2>&1
python tools/ncu_targets.py rollout4096 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2t_launches_rollout4096.csv python tools/ncu_targets.py rollout4096 > /dev/null 2>&1
ls -la gpurun_out | grep r2t
