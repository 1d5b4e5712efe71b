#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
python bench.py --decoys 16384 --resident 4096 --streams 2 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2b_c2_s2.log 2>&1; tail -c 1500 gpurun_out/r2b_c2_s2.log
python bench.py --decoys 16384 --resident 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2b_c2_r8192.log 2>&1; tail -c 1500 gpurun_out/r2b_c2_r8192.log
python bench.py --config 4 --targets 16 --steps 1 --warmup 1 > gpurun_out/r2b_c4.log 2>&1; tail -c 1500 gpurun_out/r2b_c4.log
python bench.py --config 4 --targets 16 --steps 1 --warmup 1 --streams 8 > gpurun_out/r2b_c4_s8.log 2>&1; tail -c 1500 gpurun_out/r2b_c4_s8.log
python bench.py --config 1 --streams 1 --decoys 2048 --resident 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_c1_q.log 2>&1; tail -c 1500 gpurun_out/r2b_c1_q.log
