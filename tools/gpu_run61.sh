#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/kcycle_probe.py gpu 8192 --hermitian --hermitian-setup --matrix-free --restart 8 --profile > gpurun_out/r6o_k8192.txt 2>&1
grep -E "PROFILE|second_solve" gpurun_out/r6o_k8192.txt | cut -c1-400 | head -40
