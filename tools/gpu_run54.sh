#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_host_gpu.py tests/test_shard_gpu.py -m gpu -q -x -k "matrix_free" 2>&1 | tail -5
QMG_MF_TILE=0 timeout 200 python tools/mf_probe.py 8192 | grep matrix-free | sed 's/^/simple /'
timeout 200 python tools/mf_probe.py 8192 | grep matrix-free | sed 's/^/tile   /'
