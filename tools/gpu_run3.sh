#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "hermitian or residual or link_compressed or fused_kcycle or loopback" > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3c_pytest.log
tail -5 gpurun_out/r3c_pytest.log
for t in 1 3; do QMG_TILE=$t python tools/tile_probe.py; done > gpurun_out/r3c_tile.log 2>&1; cat gpurun_out/r3c_tile.log
