#!/bin/bash
set -x
mkdir -p gpurun_out
for t in 1 3; do QMG_TILE=$t python tools/tile_probe.py; done > gpurun_out/r3d_tile.log 2>&1; cat gpurun_out/r3d_tile.log
python tools/tile_probe.py > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stencil_tma_kernel -s 4 -c 1 -o gpurun_out/r3d_tma python tools/tile_probe.py > gpurun_out/r3d_tma_ncu.log 2>&1
