#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r5i_bench1.json 2> gpurun_out/r5i_bench1.err; echo "rc $?"; tail -2 gpurun_out/r5i_bench1.err
