#!/bin/bash
# the destroy-order test from another working directory
mkdir -p gpurun_out
cd /tmp && timeout 300 python -m pytest /root/repo/tests/test_fold_gpu.py -m gpu -q -k "destroyed_before or recycled" -p no:cacheprovider > /root/repo/gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 /root/repo/gpurun_out/r2x_pytest.log
