#!/bin/bash
# round 2, call 5e: staged rbf_queries_kernel with pre-scaled coordinates: timing, full GPU suite, bench line
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rbf_queries_kernel -c 6 --csv --log-file gpurun_out/r5e_launches_rbf.csv python tools/ncu_targets.py rbf 6 > /dev/null 2>&1
grep rbf_queries gpurun_out/r5e_launches_rbf.csv | awk -F, '{print $NF}' | tr '\n' ' '; echo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r5e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5e_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r5e_bench.json 2> gpurun_out/r5e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r5e_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
