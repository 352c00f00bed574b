#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/kcycle_probe.py gpu 4096 8192 --hermitian --restart 8 --profile > gpurun_out/r3y_k.txt 2>&1; grep -v gpurun gpurun_out/r3y_k.txt | cut -c1-400
