#!/bin/bash
# round 2, GPU call A: parity suite, then the bench configurations (exploratory step counts)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --decoys 4096 --resident 4096 --steps 2 --warmup 2 --no-k1-standalone > gpurun_out/r2a_c2_4096.log 2>&1; tail -c 3000 gpurun_out/r2a_c2_4096.log
python bench.py --decoys 16384 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_c2_16384.log 2>&1; tail -c 4000 gpurun_out/r2a_c2_16384.log
python bench.py --config 1 --steps 3 --warmup 2 > gpurun_out/r2a_c1.log 2>&1; tail -c 2500 gpurun_out/r2a_c1.log
python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_c3.log 2>&1; tail -c 2500 gpurun_out/r2a_c3.log
python bench.py --config 4 --targets 16 --steps 1 --warmup 1 > gpurun_out/r2a_c4.log 2>&1; tail -c 2500 gpurun_out/r2a_c4.log
