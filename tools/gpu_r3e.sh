#!/bin/bash
# round 2, call 3e: block append -- tiled V = P Q kernel, block-per-pair Gram kernel: parity, launch list at N=4096
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mission.py -m gpu -q -x -k "append or mission or golden or duplicate or big_block or parallel or replic" > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3e_pytest.log
python tools/ncu_targets.py append > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r3e_launches_append_4096.csv python tools/ncu_targets.py append 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r3e_launches_append_4096.csv > gpurun_out/r3e_append_4096_launch_summary.txt; grep -i "append\|gemv\|rbf_cross" gpurun_out/r3e_append_4096_launch_summary.txt
