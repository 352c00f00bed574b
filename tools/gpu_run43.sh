#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "generic_stencil or dwf or fill_and_apply" 2>&1 | tail -4
PROBE_GENERIC_ONLY=1 QMG_LIB_OVERRIDE=tools/_old/libqmg_b200.so timeout 300 python tools/kernel_probe.py --only stencil --reps 10 2>&1 | sed 's/^/OLD /'
PROBE_GENERIC_ONLY=1 timeout 300 python tools/kernel_probe.py --only stencil --reps 10 2>&1 | sed 's/^/NEW /'
