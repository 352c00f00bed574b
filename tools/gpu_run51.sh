#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --scaling strong --no-cpu --parity-L 512 > gpurun_out/r6a_bench8_strong.json 2> gpurun_out/r6a_bench8_strong.err; echo "rc $?"; tail -3 gpurun_out/r6a_bench8_strong.err
