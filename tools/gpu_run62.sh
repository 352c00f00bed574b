#!/bin/bash
timeout 600 python -m pytest tests/test_shard_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "loopback or matrix_free" 2>&1 | tail -4
