#!/bin/bash
# K1 packed-f32x2 kernel: parity tests, standalone A/B against the one-decoy-per-lane kernel, bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_restraints_gpu.py tests/test_properties_gpu.py tests/test_fold_gpu.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
for d in "" "--dense"; do
  python tools/k1_bench.py $d > gpurun_out/r2c_k1_x2$d.log 2>&1; tail -1 gpurun_out/r2c_k1_x2$d.log
  TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2c_k1_scalar$d.log 2>&1; tail -1 gpurun_out/r2c_k1_scalar$d.log
  TRX_K1_CARVEOUT=100 python tools/k1_bench.py $d > gpurun_out/r2c_k1_x2_c100$d.log 2>&1; tail -1 gpurun_out/r2c_k1_x2_c100$d.log
done
python bench.py --decoys 16384 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2c_c2.log 2>&1; tail -c 1200 gpurun_out/r2c_c2.log
