#!/bin/bash
# Round 2 final bench lines on the last build (the ncu captures come from tools/gpu_round2_final.sh)
# ncu launch list of the bench command, ncu --set full captures of the top kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_default.log 2>gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference.log 2>&1; tail -c 400 gpurun_out/r2_bench_reference.log
python bench.py --decoys 4096 --resident 4096 --no-k1-standalone --no-cpu-baseline > gpurun_out/r2_bench_4096.log 2>&1; tail -c 300 gpurun_out/r2_bench_4096.log
python bench.py --config 1 > gpurun_out/r2_bench_c1.log 2>&1; tail -c 300 gpurun_out/r2_bench_c1.log
python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3.log 2>&1; tail -c 300 gpurun_out/r2_bench_c3.log
python bench.py --config 4 --steps 1 --warmup 1 --streams 8 > gpurun_out/r2_bench_c4.log 2>&1; tail -c 300 gpurun_out/r2_bench_c4.log
