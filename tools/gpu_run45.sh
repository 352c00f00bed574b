#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -4
for m in 0 1 0 1; do QMG_ROW2=$m timeout 200 python tools/rbj_probe.py 2 8192 2>&1 | sed "s/^/ROW2=$m /"; done
