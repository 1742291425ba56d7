#!/bin/bash
# round 2, call j: bench on 2 GPUs (the multi-GPU rows: replicated append in the step, config 5 strong, root-parallel search) + the 2-rank parity test
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench2 rc=$?"
grep -v "^W1\|^\*\*\*\|^$" gpurun_out/r2j_bench_2gpu.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_2gpu.json')); print({k:v for k,v in d.items() if k not in ('extra','config','roofline','clocks','cpu_baseline')})"
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2j_bench_ref_2gpu.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2j_bench_ref_2gpu.json
