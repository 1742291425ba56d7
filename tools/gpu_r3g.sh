#!/bin/bash
# round 2, call 3g (2 GPUs): result equality on two GPUs (tests/test_gpu_multi.py), 2-GPU bench line, pure-write bandwidth
set -x
mkdir -p gpurun_out
python tools/measure_write_bw.py 2>&1 | tee gpurun_out/r3g_write_bw.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r3g_pytest_multi.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3g_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3g_bench_2gpu.json 2> gpurun_out/r3g_bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r3g_bench_2gpu.err
