#!/bin/bash
# round 2, call 3n (8 GPUs): bench line at N=8 with the kernels of this session, 2-rank result-equality test
set -x
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r3n_bench_8gpu.json 2> gpurun_out/r3n_bench_8gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r3n_bench_8gpu.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r3n_bench_4gpu.json 2> gpurun_out/r3n_bench_4gpu.err; echo "bench4 rc=$?"
