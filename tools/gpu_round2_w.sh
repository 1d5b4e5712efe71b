#!/bin/bash
# configs[4] (batch mode) with the block pool: 64 targets' tables and fold batches are created and destroyed per step
mkdir -p gpurun_out
timeout 400 python bench.py --config 4 --streams 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_c4_pool.log 2> gpurun_out/r2_bench_c4_pool.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2_bench_c4_pool.log
