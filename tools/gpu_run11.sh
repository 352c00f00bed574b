#!/bin/bash
set -x
mkdir -p gpurun_out
QMG_BENCH_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --scaling strong --steps 5 --parity-L 0 > gpurun_out/r3k_bench8_strong.json 2> gpurun_out/r3k_bench8_strong.err; echo "rc $?"; grep -E "QMG-PROFILE|bench\]" gpurun_out/r3k_bench8_strong.err | head -40
