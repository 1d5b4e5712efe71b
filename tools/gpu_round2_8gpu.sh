#!/bin/bash
# Round 2, one 8xB200 box: configs[2] weak (default) and strong (4096 decoys in total), configs[3], configs[4]
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_8gpu_c2_weak.log 2>&1; tail -c 400 gpurun_out/r2_8gpu_c2_weak.log
$TR bench.py --gpus 8 --scaling strong --decoys 4096 --steps 3 --warmup 3 > gpurun_out/r2_8gpu_c2_strong.log 2>&1; tail -c 400 gpurun_out/r2_8gpu_c2_strong.log
$TR bench.py --gpus 8 --config 3 --scaling strong --steps 3 --warmup 3 > gpurun_out/r2_8gpu_c3.log 2>&1; tail -c 400 gpurun_out/r2_8gpu_c3.log
$TR bench.py --gpus 8 --config 4 --steps 2 --warmup 1 --streams 8 > gpurun_out/r2_8gpu_c4.log 2>&1; tail -c 400 gpurun_out/r2_8gpu_c4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --scaling strong --decoys 4096 --steps 3 --warmup 3 > gpurun_out/r2_2gpu_c2_strong.log 2>&1; tail -c 300 gpurun_out/r2_2gpu_c2_strong.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --scaling strong --decoys 4096 --steps 3 --warmup 3 > gpurun_out/r2_4gpu_c2_strong.log 2>&1; tail -c 300 gpurun_out/r2_4gpu_c2_strong.log
