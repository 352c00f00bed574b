#!/bin/bash
for d in 0 16 32 48 64 0 32 17 33 49; do
echo -n "dbg=$d  "; QMG_RING_DEBUG=$d QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"
done
