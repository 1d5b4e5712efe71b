#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/cart_rebuild_cost.py > gpurun_out/r2_cart_rebuild_cost.log 2>&1; cat gpurun_out/r2_cart_rebuild_cost.log
CMD="python bench.py --decoys 2048 --resident 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-k1-standalone"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60000 -c 4000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
