#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r5g_pytest.log 2>&1; tail -4 gpurun_out/r5g_pytest.log
for t in 9 3; do QMG_TILE=$t TILE_PROBE_SMALL=1 TILE_PROBE_ISOLATED=1 timeout 120 python tools/tile_probe.py 2>&1 | grep -A2 "herm=1"; done
python bench.py > gpurun_out/r5g_bench1.json 2> gpurun_out/r5g_bench1.err; echo "rc $?"; tail -2 gpurun_out/r5g_bench1.err
