#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fine_level_link or hermitian or link_compressed or batched_qr or loopback" 2>&1 | tail -4
python tools/kernel_probe.py --only stencil --reps 10 2>&1 | grep -A1 "nc=2"
