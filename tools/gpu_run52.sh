#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_host_gpu.py tests/test_shard_gpu.py -m gpu -q -x -k "matrix_free" 2>&1 | tail -15
