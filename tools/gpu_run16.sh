#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tile_kernel_flavours" > gpurun_out/r3o_pytest.log 2>&1; tail -5 gpurun_out/r3o_pytest.log
for t in 1 7 8; do QMG_TILE=$t timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"; done > gpurun_out/r3o_tile.log 2>&1; cat gpurun_out/r3o_tile.log
