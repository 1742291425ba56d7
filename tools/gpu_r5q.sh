#!/bin/bash
# round 2, call 5q: whole-stage / all-rows fast path in the DMMA tile loader -- full GPU suite, bench line, refresh time
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r5q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5q_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r5q_bench.json 2> gpurun_out/r5q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r5q_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])
        for k in ('config4_grid_stage_ms','config4_grid_stage_general_grid_ms','config5_reward_evals_per_s','mcts_rollouts_per_s','replicate_recompute_ms'): print(k, d.get(k))
PY
