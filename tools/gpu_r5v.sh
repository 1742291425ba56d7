#!/bin/bash
# round 2, call 5v: the tree as committed at the end of the round -- smoke, full GPU suite, default bench line
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r5v_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r5v_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r5v_pytest.log
timeout 900 python bench.py > gpurun_out/r5v_bench.json 2> gpurun_out/r5v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r5v_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic_source'], d['clocks'])
PY
