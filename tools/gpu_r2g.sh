#!/bin/bash
# round 2, call g: rollout v5 (RED limbs), potrf 2D cyclic, full suite, bench at N=1
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2g_pytest.log
timeout 300 python tools/mcts_split.py > gpurun_out/r2g_split.log 2>&1; cat gpurun_out/r2g_split.log
python tools/ncu_targets.py factor > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2g_launches_factor4096.csv python tools/ncu_targets.py factor 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2g_launches_factor4096.csv > gpurun_out/r2g_factor_summary.txt; head -4 gpurun_out/r2g_factor_summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print({k:v for k,v in d.items() if k not in ('extra','config','roofline','clocks','cpu_baseline')})"
