#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_restraints_gpu.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
for d in "" "--dense"; do
  python tools/k1_bench.py $d > gpurun_out/r2g_default$d.log 2>&1; echo "default(stepped scalar st1 c72) $d: $(tail -1 gpurun_out/r2g_default$d.log)"
  TRX_K1_CARVEOUT=86 python tools/k1_bench.py $d > gpurun_out/r2g_default_c86$d.log 2>&1; echo "default c86 $d: $(tail -1 gpurun_out/r2g_default_c86$d.log)"
  TRX2DYN_LIB=$PWD/trrosettax2-dynamics_b200/libtrx2dyn_noalloc.so python tools/k1_bench.py $d > gpurun_out/r2g_noalloc$d.log 2>&1; echo "noalloc $d: $(tail -1 gpurun_out/r2g_noalloc$d.log)"
  TRX_K1_STAGE=0 python tools/k1_bench.py $d > gpurun_out/r2g_st0$d.log 2>&1; echo "stage0 $d: $(tail -1 gpurun_out/r2g_st0$d.log)"
  TRX_K1_SYM=1 python tools/k1_bench.py $d > gpurun_out/r2g_sym$d.log 2>&1; echo "sym $d: $(tail -1 gpurun_out/r2g_sym$d.log)"
  TRX_K1_FREE=1 TRX_K1_SYM=1 python tools/k1_bench.py $d > gpurun_out/r2g_free_sym$d.log 2>&1; echo "free sym $d: $(tail -1 gpurun_out/r2g_free_sym$d.log)"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
