#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3b_pytest.log
tail -15 gpurun_out/r3b_pytest.log
python tools/kernel_probe.py --only transfer > gpurun_out/r3b_transfer.log 2>&1; cat gpurun_out/r3b_transfer.log
python tools/kcycle_probe.py gpu 4096 --hermitian --restart 8 --profile > gpurun_out/r3b_k4096.txt 2>&1; grep -v gpurun gpurun_out/r3b_k4096.txt | head -12
