#!/bin/bash
# round 2, call 3m: full GPU suite after the session's kernel changes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3m_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r3m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3m_smoke.log
