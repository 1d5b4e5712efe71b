#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_fold_gpu.py tests/test_properties_gpu.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k1-standalone > gpurun_out/r2m_c2.log 2>&1; tail -c 300 gpurun_out/r2m_c2.log
python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_c3.log 2>&1; tail -c 300 gpurun_out/r2m_c3.log
