#!/bin/bash
# K1 side-packed f32x2 kernel + bulk-staged pair records + L-BFGS bulk-async ring: parity tests, A/B, bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_restraints_gpu.py tests/test_properties_gpu.py tests/test_fold_gpu.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
for d in "" "--dense"; do
  python tools/k1_bench.py $d > gpurun_out/r2d_k1_sym$d.log 2>&1; tail -1 gpurun_out/r2d_k1_sym$d.log
  TRX_K1_CARVEOUT=86 python tools/k1_bench.py $d > gpurun_out/r2d_k1_sym_c86$d.log 2>&1; tail -1 gpurun_out/r2d_k1_sym_c86$d.log
  TRX_K1_NO_STAGE=1 python tools/k1_bench.py $d > gpurun_out/r2d_k1_sym_nostage$d.log 2>&1; tail -1 gpurun_out/r2d_k1_sym_nostage$d.log
  TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2d_k1_scalar$d.log 2>&1; tail -1 gpurun_out/r2d_k1_scalar$d.log
  TRX_K1_SCALAR=1 TRX_K1_NO_STAGE=1 python tools/k1_bench.py $d > gpurun_out/r2d_k1_scalar_nostage$d.log 2>&1; tail -1 gpurun_out/r2d_k1_scalar_nostage$d.log
done
python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2d_c2.log 2>&1; tail -c 1200 gpurun_out/r2d_c2.log
TRX_NO_LB_RING=1 python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2d_c2_noring.log 2>&1; tail -c 1200 gpurun_out/r2d_c2_noring.log
