#!/bin/bash
mkdir -p gpurun_out
python tools/dynamics_profile.py > gpurun_out/r2_dynamics_profile.log 2>&1; cat gpurun_out/r2_dynamics_profile.log
