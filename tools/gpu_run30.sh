#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tile_kernel_flavours or gcr_orthogonalize or krylov_step or residual_epilogue or batched_qr or in_place" > gpurun_out/r4c_plain.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tile_kernel_flavours or gcr_orthogonalize or krylov_step or residual_epilogue or batched_qr or in_place" > gpurun_out/r4c_memcheck.log 2>&1; echo "memcheck rc $?"; tail -12 gpurun_out/r4c_memcheck.log
