#!/bin/bash
# last build: full GPU suite, default bench line (driver flags would be --steps 20 --warmup 5; here 3/3), configs[1], single batch
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_default.log 2>gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.log
python bench.py --decoys 4096 --resident 4096 --no-k1-standalone --no-cpu-baseline > gpurun_out/r2_bench_4096.log 2>&1; tail -c 300 gpurun_out/r2_bench_4096.log
python bench.py --config 1 > gpurun_out/r2_bench_c1.log 2>&1; tail -c 300 gpurun_out/r2_bench_c1.log
python bench.py --decoys 512 --resident 512 --no-k1-standalone --no-cpu-baseline > gpurun_out/r2_bench_512.log 2>&1; tail -c 300 gpurun_out/r2_bench_512.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
