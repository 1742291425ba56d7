#!/bin/bash
# round 2, call 4d: i8x6 with the A slices staged in tensor memory (tcgen05.cp + TS-mode MMA) -- correctness and speed against the SS form
set -x
mkdir -p gpurun_out
for n in 1000 4096; do
  IPP_I8_TS=0 timeout 120 python tools/check_i8_ts.py $n 2>&1 | tail -2 | tee -a gpurun_out/r4d_i8_ts.log
  IPP_I8_TS=1 timeout 120 python tools/check_i8_ts.py $n 2>&1 | tail -2 | tee -a gpurun_out/r4d_i8_ts.log
done
nvidia-smi --query-gpu=name,temperature.gpu --format=csv | tail -1
