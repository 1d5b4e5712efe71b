#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_fold_gpu.py -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2k_c2.log 2>&1; tail -c 1000 gpurun_out/r2k_c2.log
