#!/bin/bash
set -x
mkdir -p gpurun_out
QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1" &&
QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:stencil_ring_kernel -s 3 -c 1 -o gpurun_out/r5f_ring python tools/tile_probe.py > gpurun_out/r5f_ring_ncu.log 2>&1
ls -la gpurun_out/
