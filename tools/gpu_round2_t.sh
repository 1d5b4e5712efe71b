#!/bin/bash
# final validation of the build with the block pool and the 12288-position default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2t_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_default12k.log 2> gpurun_out/r2_bench_default12k.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_default12k.log
