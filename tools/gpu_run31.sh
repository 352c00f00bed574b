#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tile_kernel_flavours or residual_epilogue" 2>&1 | tail -3
for t in 9 3 9 3; do QMG_TILE=$t TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"; done
QMG_RING_DEBUG=2 QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"
QMG_RING_DEBUG=1 QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"
