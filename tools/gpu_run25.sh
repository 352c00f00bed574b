#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/latency_probe.py 2>&1 | tail -16
python tools/kernel_probe.py --only blas 2>&1 | tail -19
python -m pytest tests -m gpu -q > gpurun_out/r3x_pytest.log 2>&1; tail -8 gpurun_out/r3x_pytest.log
