#!/bin/bash
# context reference counting + two bench variants of the default workload (2 streams; 8192 resident positions)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fold_gpu.py -m gpu -x -q -k "recycled or destroyed_before or continuous" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest.log
python tools/dynamics_profile.py > gpurun_out/r2r_dynamics_profile.log 2>&1; echo "profile rc=$?"; tail -3 gpurun_out/r2r_dynamics_profile.log
timeout 600 python bench.py --steps 2 --warmup 1 --streams 2 --no-k1-standalone > gpurun_out/r2r_bench_s2.json 2> gpurun_out/r2r_bench_s2.err; echo "bench s2 rc=$?"; cut -c1-400 gpurun_out/r2r_bench_s2.json
timeout 600 python bench.py --steps 2 --warmup 1 --resident 8192 --no-k1-standalone > gpurun_out/r2r_bench_r8192.json 2> gpurun_out/r2r_bench_r8192.err; echo "bench r8192 rc=$?"; cut -c1-400 gpurun_out/r2r_bench_r8192.json
