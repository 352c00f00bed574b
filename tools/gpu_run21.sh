#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tile_kernel_flavours" 2>&1 | tail -3
for t in 12 9 3; do QMG_TILE=$t TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"; done
for t in 12 9 3; do QMG_TILE=$t TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"; done
