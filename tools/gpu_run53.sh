#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/mf_probe.py 8192
timeout 200 python tools/mf_probe.py 1024
python bench.py > gpurun_out/r6c_bench1.json 2> gpurun_out/r6c_bench1.err; echo "rc $?"; tail -2 gpurun_out/r6c_bench1.err
