#!/bin/bash
# round 2, call b: the restructured bench at N=1 and N=2
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2b_bench_2gpu.err
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
