#!/bin/bash
# 2 x B200 sanity of the final build: default workload, weak scaling, one timed step
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 1 --warmup 1 --no-k1-standalone > gpurun_out/r2_2gpu_c2_weak_final.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/r2_2gpu_c2_weak_final.log | cut -c1-600
