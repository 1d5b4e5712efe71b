#!/bin/bash
# last checks: the fold tests of the last build, the default line as the driver will run it (shorter), two data points
set -x
mkdir -p gpurun_out
python -m pytest tests/test_fold_gpu.py -m gpu -q > gpurun_out/r2_pytest_fold_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_fold_last.log; tail -3 gpurun_out/r2_pytest_fold_last.log
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2_bench_default24k.log 2> gpurun_out/r2_bench_default24k.err; tail -c 400 gpurun_out/r2_bench_default24k.log; tail -4 gpurun_out/r2_bench_default24k.err
python bench.py --decoys 65536 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2_bench_65536.log 2>&1; tail -c 300 gpurun_out/r2_bench_65536.log
python bench.py --streams 2 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2_bench_s2.log 2>&1; tail -c 300 gpurun_out/r2_bench_s2.log
