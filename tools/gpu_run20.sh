#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tile_kernel_flavours" 2>&1 | tail -3
QMG_TILE=9 python tools/kcycle_probe.py gpu 8192 --hermitian --restart 8 > gpurun_out/r3s_k8192_ring.txt 2>&1; tail -1 gpurun_out/r3s_k8192_ring.txt | cut -c1-330
QMG_TILE=1 python tools/kcycle_probe.py gpu 8192 --hermitian --restart 8 > gpurun_out/r3s_k8192_cpasync.txt 2>&1; tail -1 gpurun_out/r3s_k8192_cpasync.txt | cut -c1-330
