#!/bin/bash
set -x
mkdir -p gpurun_out
for t in 1 3; do QMG_TILE=$t python tools/tile_probe.py; done > gpurun_out/r3f_tile.log 2>&1; cat gpurun_out/r3f_tile.log
python tools/kernel_probe.py --only blas > gpurun_out/r3f_blas.log 2>&1; grep multi gpurun_out/r3f_blas.log
python -m pytest tests -m gpu -q -k "hermitian or residual or link_compressed or fused_kcycle or loopback or blas or solvers" > gpurun_out/r3f_pytest.log 2>&1; tail -3 gpurun_out/r3f_pytest.log
python tools/kcycle_probe.py gpu 4096 --hermitian --hermitian-setup --restart 8 > gpurun_out/r3f_k4096_hsetup.txt 2>&1; tail -1 gpurun_out/r3f_k4096_hsetup.txt | cut -c1-330
