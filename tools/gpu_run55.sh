#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r6g_pytest.log 2>&1; tail -4 gpurun_out/r6g_pytest.log
python bench.py > gpurun_out/r6g_bench1.json 2> gpurun_out/r6g_bench1.err; echo "rc $?"; tail -2 gpurun_out/r6g_bench1.err
