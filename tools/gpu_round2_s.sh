#!/bin/bash
# resident positions of the default workload: 12288 and 16384 (8192: 1625 decoys/s, 4096: 1563)
mkdir -p gpurun_out
for r in 12288 16384; do
timeout 600 python bench.py --steps 2 --warmup 1 --resident $r --no-k1-standalone > gpurun_out/r2s_bench_r$r.json 2> gpurun_out/r2s_bench_r$r.err; echo "bench r$r rc=$?"; cut -c1-200 gpurun_out/r2s_bench_r$r.json
done
