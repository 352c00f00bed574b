#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -q -x 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 > gpurun_out/r5k_bench2.json 2> gpurun_out/r5k_bench2.err; echo "rc $?"; tail -3 gpurun_out/r5k_bench2.err
