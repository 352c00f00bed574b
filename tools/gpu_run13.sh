#!/bin/bash
set -x
mkdir -p gpurun_out
for t in 1 4 5 6; do QMG_TILE=$t python tools/tile_probe.py 2>&1 | grep "herm=1"; done > gpurun_out/r3l_tile.log 2>&1; cat gpurun_out/r3l_tile.log
python -m pytest tests -m gpu -q -x > gpurun_out/r3l_pytest.log 2>&1; tail -5 gpurun_out/r3l_pytest.log
python tools/kcycle_probe.py gpu 4096 8192 --hermitian --restart 8 --profile > gpurun_out/r3l_k.txt 2>&1; grep -v gpurun gpurun_out/r3l_k.txt | cut -c1-330
