#!/bin/bash
set -x
mkdir -p gpurun_out
for t in 1 4 5 6; do QMG_TILE=$t python tools/tile_probe.py 2>&1 | grep "herm=1"; done > gpurun_out/r3l_tile.log 2>&1; cat gpurun_out/r3l_tile.log
