#!/bin/bash
# usage: tools/build_variant.sh NAME "-DIPP_BK=32 -DIPP_STAGES=3"   -> csrc/libipp_b200_NAME.so (kernel-variant experiments)
set -e
cd "$(dirname "$0")/../informative_path_planning_b200/csrc"
out=/tmp/ippvar_$1; mkdir -p $out
for f in common rbf gemm predict factor append rewards rff infogain paths rollout predict_i8 tree; do
  nvcc -O3 -std=c++20 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $2 -c $f.cu -o $out/$f.o &
done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libipp_b200_$1.so $out/*.o
