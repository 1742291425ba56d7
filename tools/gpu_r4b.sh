#!/bin/bash
# round 2, call 4b: final bench line and full suite after the last kernel changes (small-tile GEMM, batched weight draws)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4b_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r4b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r4b_bench.json 2> gpurun_out/r4b_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r4b_bench.err
