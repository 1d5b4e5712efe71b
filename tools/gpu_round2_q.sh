#!/bin/bash
# block pool validation: GPU tests, the dynamics-loop profile, a short default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
python tools/dynamics_profile.py > gpurun_out/r2q_dynamics_profile.log 2>&1; cat gpurun_out/r2q_dynamics_profile.log
timeout 600 python tools/dynamics_example.py --n-max 300 > gpurun_out/r2q_dynamics_example.log 2>&1; tail -12 gpurun_out/r2q_dynamics_example.log
timeout 900 python bench.py --steps 3 --warmup 1 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; cat gpurun_out/r2q_bench.json
