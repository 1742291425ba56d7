#!/bin/bash
# round 2, call 4g (2 GPUs): bench lines at N=2 and N=1 with the sharded config 4 row
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r4g_bench_2gpu.json 2> gpurun_out/r4g_bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r4g_bench_2gpu.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r4g_bench_1gpu.json 2> gpurun_out/r4g_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for f in ('r4g_bench_1gpu.json', 'r4g_bench_2gpu.json'):
    for l in open('gpurun_out/' + f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, d['n_gpus'], d['value'], {k: v for k, v in d.items() if k.startswith('config4')})
PY
