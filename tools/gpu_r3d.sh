#!/bin/bash
# round 2, call 3d: sampler pipelining test, bench line, launch lists (FP64 and i8x6 predict)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sample_max_vals" > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3d_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3d_bench.err
python bench.py --steps 2 --warmup 1 --no-extras --queries 131072 --precision i8x6 > gpurun_out/r3d_plain_i8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3d_launches_bench_i8x6_q131072.csv python bench.py --steps 2 --warmup 1 --no-extras --queries 131072 --precision i8x6 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3d_launches_bench_i8x6_q131072.csv | head -8
