#!/bin/bash
# round 2, call 3o: BASELINE configs 1 and 2 as whole missions with the final kernels
set -x
mkdir -p gpurun_out
timeout 1500 python tools/bench_mission.py > gpurun_out/r3o_mission_bench.log 2>&1; echo "rc=$?"
cp gpurun_out/mission_bench.json gpurun_out/r3o_mission_bench.json
python - <<'PY'
import json
for m in json.load(open('gpurun_out/r3o_mission_bench.json')):
    print(m['kind'], 'noise', m['noise'], 'steps', m['steps_done'], 'mission_s %.2f' % m['mission_s'], 'median_step_ms %.2f' % m['median_step_ms'],
          'median_plan_ms %.2f' % m['median_plan_ms'], 'n_obs_final', m['n_obs_final'], [(c['t'], round(c['cpu_step_s'], 2), round(c['device_step_s'], 4)) for c in m['cpu_snapshots']])
PY
