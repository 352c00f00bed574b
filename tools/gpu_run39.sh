#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/setup_profile.py 8192 > gpurun_out/r5j_setup8192.txt 2>&1; head -30 gpurun_out/r5j_setup8192.txt
