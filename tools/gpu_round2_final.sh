#!/bin/bash
# Round 2 final single-GPU evidence: default bench line, single-batch reading, reference arm, configs 1/3/4,
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
# ---- ncu (each command has just exited 0 without ncu: the plain runs below)
CMD="python bench.py --decoys 2048 --resident 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-k1-standalone"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60000 -c 4000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
python tools/k1_bench.py --iters 3 > gpurun_out/r2_k1_plain_sparse.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:restraints_kernel -s 3 -c 1 -o gpurun_out/r2_k1_sparse python tools/k1_bench.py --iters 3 > gpurun_out/r2_ncu_k1_sparse.log 2>&1
python tools/k1_bench.py --iters 3 --dense > gpurun_out/r2_k1_plain_dense.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:restraints_kernel -s 3 -c 1 -o gpurun_out/r2_k1_dense python tools/k1_bench.py --iters 3 --dense > gpurun_out/r2_ncu_k1_dense.log 2>&1
for k in vdw_kernel lbfgs_dots_ring lbfgs_update_ring lbfgs_step nerf_kernel torsion_grad; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 400 -c 1 -o gpurun_out/r2_$k $CMD > gpurun_out/r2_ncu_$k.log 2>&1
done
python tools/dynamics_example.py --n-max 40 > gpurun_out/r2_dynamics_example.log 2>&1; cat gpurun_out/r2_dynamics_example.log
ls -la gpurun_out/*.ncu-rep
