#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_gpu.py -m gpu -q -x -k "bicgstab or solvers_iteration or n13_kcycle or n22 or native_kcycle" 2>&1 | tail -6
QMG_BICGSTAB_FUSED=0 timeout 300 python tools/setup_profile.py 4096 > gpurun_out/r5h_setup_unfused.txt 2>&1; head -12 gpurun_out/r5h_setup_unfused.txt
QMG_BICGSTAB_FUSED=1 timeout 300 python tools/setup_profile.py 4096 > gpurun_out/r5h_setup_fused.txt 2>&1; head -14 gpurun_out/r5h_setup_fused.txt
