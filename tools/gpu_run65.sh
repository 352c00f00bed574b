#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -q -x -k "bicgstab or n13_kcycle_parity" 2>&1 | tail -3
