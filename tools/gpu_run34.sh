#!/bin/bash
for d in 0 20 36 52 24 40 56 0 36 40 37 41; do
echo -n "dbg=$d  "; QMG_RING_DEBUG=$d QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1"
done
