#!/bin/bash
# default workload as two lanes of 6144 positions on two streams
mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 1 --streams 2 --no-k1-standalone --no-cpu-baseline > gpurun_out/r2_bench_s2_12k.log 2> gpurun_out/r2_bench_s2_12k.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2_bench_s2_12k.log
