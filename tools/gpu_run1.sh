#!/bin/bash
# round 2, GPU call 1: parity suite, fused vs unfused K-cycle, ncu of the kernels VERDICT names
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3a_pytest.log
tail -5 gpurun_out/r3a_pytest.log
python tools/kcycle_probe.py gpu 4096 --hermitian --restart 8 --unfused --profile > gpurun_out/r3a_k4096_unfused.txt 2>&1
python tools/kcycle_probe.py gpu 4096 8192 --hermitian --restart 8 --profile > gpurun_out/r3a_k_fused.txt 2>&1
tail -3 gpurun_out/r3a_k4096_unfused.txt gpurun_out/r3a_k_fused.txt
python tools/tile_probe.py > gpurun_out/r3a_tile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stencil_tile_kernel -s 4 -c 2 -o gpurun_out/r3a_tile python tools/tile_probe.py > gpurun_out/r3a_tile_ncu.log 2>&1
python tools/kernel_probe.py --only transfer --reps 3 > gpurun_out/r3a_transfer_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:restrict_kernel -s 3 -c 2 -o gpurun_out/r3a_restrict python tools/kernel_probe.py --only transfer --reps 3 > gpurun_out/r3a_restrict_ncu.log 2>&1
python tools/kernel_probe.py --only blas --reps 3 > gpurun_out/r3a_blas_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:reduce_kernel -s 3 -c 4 -o gpurun_out/r3a_reduce python tools/kernel_probe.py --only blas --reps 3 > gpurun_out/r3a_reduce_ncu.log 2>&1
cat gpurun_out/r3a_tile_plain.log gpurun_out/r3a_transfer_plain.log
