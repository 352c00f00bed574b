#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/latency_probe.py > gpurun_out/r3n_latency1.txt 2>&1; cat gpurun_out/r3n_latency1.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/latency_probe.py > gpurun_out/r3n_latency8.txt 2>&1; grep -v "^\*\|OMP\|NCCL" gpurun_out/r3n_latency8.txt
