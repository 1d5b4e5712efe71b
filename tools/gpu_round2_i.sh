#!/bin/bash
# Verlet list of the vdw / hydrogen-bond pair search: bit-identity test, fold tests, A/B on configs 2 and 3
set -x
mkdir -p gpurun_out
python -m pytest tests/test_fold_gpu.py tests/test_properties_gpu.py tests/test_configs_gpu.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -12 gpurun_out/r2i_pytest.log
python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2i_c2.log 2>&1; tail -c 1300 gpurun_out/r2i_c2.log
TRX_NO_NBL=1 python bench.py --decoys 8192 --resident 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2i_c2_nonbl.log 2>&1; tail -c 1300 gpurun_out/r2i_c2_nonbl.log
python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_c3.log 2>&1; tail -c 1300 gpurun_out/r2i_c3.log
TRX_NO_NBL=1 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_c3_nonbl.log 2>&1; tail -c 1300 gpurun_out/r2i_c3_nonbl.log
