#!/bin/bash
# the last build: full GPU suite, smoke, and every single-GPU bench line again
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_default24k.log 2>gpurun_out/r2_bench_default24k.err; tail -c 400 gpurun_out/r2_bench_default24k.log
python bench.py --decoys 4096 --resident 4096 --no-k1-standalone --no-cpu-baseline > gpurun_out/r2_bench_4096.log 2>&1; tail -c 300 gpurun_out/r2_bench_4096.log
python bench.py --config 1 > gpurun_out/r2_bench_c1.log 2>&1; tail -c 300 gpurun_out/r2_bench_c1.log
python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3.log 2>&1; tail -c 300 gpurun_out/r2_bench_c3.log
python bench.py --config 4 --steps 1 --warmup 1 --streams 8 > gpurun_out/r2_bench_c4.log 2>&1; tail -c 300 gpurun_out/r2_bench_c4.log
python bench.py --decoys 512 --resident 512 --no-k1-standalone --no-cpu-baseline > gpurun_out/r2_bench_512.log 2>&1; tail -c 300 gpurun_out/r2_bench_512.log
CMD="python bench.py --decoys 2048 --resident 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-k1-standalone"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vdw_kernel -s 400 -c 1 -o gpurun_out/r2_vdw_kernel $CMD > gpurun_out/r2_ncu_vdw_kernel.log 2>&1
