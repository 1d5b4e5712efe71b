#!/bin/bash
# K1 with the tile's coordinates staged in shared memory: parity tests, A/B matrix, bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_restraints_gpu.py tests/test_properties_gpu.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
for d in "" "--dense"; do
 for st in 2 1; do
  TRX_K1_STAGE=$st python tools/k1_bench.py $d > gpurun_out/r2f_free_sym_st$st$d.log 2>&1; echo "free sym stage$st $d: $(tail -1 gpurun_out/r2f_free_sym_st$st$d.log)"
  TRX_K1_STAGE=$st TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2f_free_scalar_st$st$d.log 2>&1; echo "free scalar stage$st $d: $(tail -1 gpurun_out/r2f_free_scalar_st$st$d.log)"
  TRX_K1_STAGE=$st TRX_K1_STEPPED=1 TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2f_stepped_scalar_st$st$d.log 2>&1; echo "stepped scalar stage$st $d: $(tail -1 gpurun_out/r2f_stepped_scalar_st$st$d.log)"
  TRX_K1_STAGE=$st TRX_K1_STEPPED=1 python tools/k1_bench.py $d > gpurun_out/r2f_stepped_sym_st$st$d.log 2>&1; echo "stepped sym stage$st $d: $(tail -1 gpurun_out/r2f_stepped_sym_st$st$d.log)"
 done
 TRX_K1_STAGE=1 TRX_K1_CARVEOUT=72 TRX_K1_STEPPED=1 TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2f_stepped_scalar_st1_c72$d.log 2>&1; echo "stepped scalar stage1 carve72 $d: $(tail -1 gpurun_out/r2f_stepped_scalar_st1_c72$d.log)"
 TRX_K1_STAGE=0 TRX_K1_CARVEOUT=72 TRX_K1_STEPPED=1 TRX_K1_SCALAR=1 python tools/k1_bench.py $d > gpurun_out/r2f_stepped_scalar_st0_c72$d.log 2>&1; echo "stepped scalar stage0 carve72 $d: $(tail -1 gpurun_out/r2f_stepped_scalar_st0_c72$d.log)"
done
