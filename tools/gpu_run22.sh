#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/kcycle_probe.py gpu 8192 --hermitian --restart 8 --profile > gpurun_out/r3u_k8192.txt 2>&1; grep -v gpurun gpurun_out/r3u_k8192.txt | cut -c1-400
python -m pytest tests -m gpu -q > gpurun_out/r3u_pytest.log 2>&1; tail -3 gpurun_out/r3u_pytest.log
