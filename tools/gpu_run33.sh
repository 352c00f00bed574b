#!/bin/bash
set -x
for r in 1 2; do
QMG_TILE=1 TILE_PROBE_SMALL=1 QMG_LIB_OVERRIDE=tools/_old/libqmg_b200.so timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1" | sed 's/^/OLD /'
QMG_TILE=9 TILE_PROBE_SMALL=1 timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1" | sed 's/^/NEW /'
done
QMG_RING_DEBUG=2 QMG_TILE=1 TILE_PROBE_SMALL=1 QMG_LIB_OVERRIDE=tools/_old/libqmg_b200.so timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1" | sed 's/^/OLD dbg2 /'
QMG_RING_DEBUG=1 QMG_TILE=1 TILE_PROBE_SMALL=1 QMG_LIB_OVERRIDE=tools/_old/libqmg_b200.so timeout 120 python tools/tile_probe.py 2>&1 | grep "herm=1" | sed 's/^/OLD dbg1 /'
