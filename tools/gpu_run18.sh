#!/bin/bash
set -x
mkdir -p gpurun_out
for d in 0 1 2; do echo "ring debug $d"; QMG_RING_DEBUG=$d QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"; done > gpurun_out/r3q_ring.log 2>&1; cat gpurun_out/r3q_ring.log
QMG_TILE=1 TILE_PROBE_SMALL=1 python tools/tile_probe.py 2>&1 | grep "herm=1"
