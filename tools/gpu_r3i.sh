#!/bin/bash
# round 2, call 3h: meshgrid form of the grid stage -- parity, timing against the general grid
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rff or max_vals" > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3i_pytest.log
timeout 300 python tools/bench_rff.py 100 2>&1 | tee gpurun_out/r3i_bench_rff.log
python tools/ncu_targets.py rffmesh > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3i_launches_rff_mesh.csv python tools/ncu_targets.py rffmesh > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3i_launches_rff_mesh.csv > gpurun_out/r3i_rff_mesh_launch_summary.txt; head -6 gpurun_out/r3i_rff_mesh_launch_summary.txt
