#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/dynamics_example.py --n-max 300 > gpurun_out/r2_dynamics_example_nmax300.log 2>&1; cat gpurun_out/r2_dynamics_example_nmax300.log
