#!/bin/bash
# round 2, call 4e: full GPU suite on the final tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r4e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
