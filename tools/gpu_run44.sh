#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python tools/rbj_probe.py 8 2048
timeout 200 python tools/rbj_probe.py 2 8192
REPS=1 timeout 400 ncu --set full --clock-control none -k regex:stencil_kernel -s 6 -c 4 -o gpurun_out/r5q_rbj python tools/rbj_probe.py 8 2048 > gpurun_out/r5q_ncu.log 2>&1
