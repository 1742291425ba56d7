#!/bin/bash
# round 2, call 5n (2 GPUs): 2-rank equality tests and the bench line at N=2 on the final tree
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r5n_pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5n_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r5n_bench_2gpu.json 2> gpurun_out/r5n_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r5n_bench_2gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['n_gpus'], d['value'], d['e2e']['value'], d.get('config5_reward_evals_per_s'), d.get('mcts_rollouts_per_s'), d.get('config4_sharded_end_to_end_ms'))
PY
