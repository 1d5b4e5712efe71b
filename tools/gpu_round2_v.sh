#!/bin/bash
# full GPU suite after the fix of the tables-create failure path (context reference was given back twice)
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2v_smoke.log
