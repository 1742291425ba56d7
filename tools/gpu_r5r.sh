#!/bin/bash
# round 2, call 5r (8 GPUs): bench line at N=8 on the final tree of the round
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r5r_bench_8gpu.json 2> gpurun_out/r5r_bench_8gpu.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r5r_bench_8gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r5r_bench_8gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['n_gpus'], d['value'], d['e2e']['value'], d.get('config5_reward_evals_per_s'), d.get('mcts_rollouts_per_s'), d.get('config4_sharded_end_to_end_ms'), d['clocks'])
PY
