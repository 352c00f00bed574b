#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r5u_pytest.log 2>&1; tail -4 gpurun_out/r5u_pytest.log
python bench.py > gpurun_out/r5u_bench1.json 2> gpurun_out/r5u_bench1.err; echo "rc $?"; tail -2 gpurun_out/r5u_bench1.err
