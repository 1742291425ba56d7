#!/bin/bash
# round 2, call z: hotspot_info rollouts in the native search -- parity with the Python tree, rate
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mission.py -m gpu -q -x -k "mission" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2z_pytest.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2z_mcts_info_rates.log
import sys; sys.path.insert(0, 'tools')
import bench_mcts as bm
for f in ('info_gain', 'hotspot_info'):
    for native in (True, False):
        bm.run('gpu', 900, f, budget=20, native=native)
        r = min((bm.run('gpu', 900, f, native=native) for _ in range(3)), key=lambda x: x['loop_seconds'])
        print('N=900 %s %s: %.0f rollouts/s over the rollout loop, planning step %.1f ms, root visits %s' % (f, r['kind'], r['loop_rollouts_per_s'], 1e3 * r['seconds'], r['root_visits']))
PY
