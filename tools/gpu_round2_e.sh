#!/bin/bash
# barrier-free K1 (fixed-point shared atomics): parity tests, A/B against the stepped kernel, bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_restraints_gpu.py tests/test_properties_gpu.py tests/test_fold_gpu.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
for d in "" "--dense"; do
  python tools/k1_bench.py $d > gpurun_out/r2e_k1_free_sym$d.log 2>&1; tail -1 gpurun_out/r2e_k1_free_sym$d.log
  TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2e_k1_free_scalar$d.log 2>&1; tail -1 gpurun_out/r2e_k1_free_scalar$d.log
  TRX_K1_STEPPED=1 TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2e_k1_stepped_scalar$d.log 2>&1; tail -1 gpurun_out/r2e_k1_stepped_scalar$d.log
  TRX_K1_STEPPED=1 python tools/k1_bench.py $d > gpurun_out/r2e_k1_stepped_sym$d.log 2>&1; tail -1 gpurun_out/r2e_k1_stepped_sym$d.log
done
python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2e_c2.log 2>&1; tail -c 1200 gpurun_out/r2e_c2.log
