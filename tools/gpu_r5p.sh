#!/bin/bash
# round 2, call 5p: final tree of the round -- smoke, full GPU suite, bench line, reference arm, launch list of the bench
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r5p_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r5p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5p_pytest.log
timeout 900 python bench.py > gpurun_out/r5p_bench.json 2> gpurun_out/r5p_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r5p_bench_reference.json 2> gpurun_out/r5p_bench_reference.err; echo "bench ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r5p_launches_bench.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r5p_ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r5p_launches_bench.csv > gpurun_out/r5p_bench_launch_summary.txt 2>&1; head -12 gpurun_out/r5p_bench_launch_summary.txt
python - <<'PY'
import json
for f in ('r5p_bench.json', 'r5p_bench_reference.json'):
    for l in open('gpurun_out/' + f):
        if l.startswith('{'):
            d = json.loads(l); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline', {}).get('frac'), d.get('mcts_rollouts_per_s'), d.get('config4_grid_stage_general_grid_ms'))
PY
