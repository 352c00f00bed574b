#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/kcycle_probe.py gpu 8192 --hermitian --hermitian-setup --restart 8 --profile > gpurun_out/r5v_k8192.txt 2>&1
grep -E "PROFILE|second_solve" gpurun_out/r5v_k8192.txt | cut -c1-200 | head -40
