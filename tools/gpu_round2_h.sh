#!/bin/bash
# full GPU suite with the hydrogen-bond term and the device dynamics step; K1 stage 0 vs 1 at their own smem sizes; bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -15 gpurun_out/r2h_pytest.log
for d in "" "--dense"; do
  python tools/k1_bench.py $d > gpurun_out/r2h_st1$d.log 2>&1; echo "stage1 $d: $(tail -1 gpurun_out/r2h_st1$d.log)"
  TRX_K1_STAGE=0 python tools/k1_bench.py $d > gpurun_out/r2h_st0$d.log 2>&1; echo "stage0 $d: $(tail -1 gpurun_out/r2h_st0$d.log)"
done
python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2h_c2.log 2>&1; tail -c 1500 gpurun_out/r2h_c2.log
TRX_K1_STAGE=0 python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2h_c2_st0.log 2>&1; tail -c 1500 gpurun_out/r2h_c2_st0.log
