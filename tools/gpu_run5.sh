#!/bin/bash
set -x
mkdir -p gpurun_out
QMG_TILE=1 python tools/kcycle_probe.py gpu 8192 --hermitian --restart 8 > gpurun_out/r3e_k8192_tma.txt 2>&1; tail -1 gpurun_out/r3e_k8192_tma.txt | cut -c1-400
QMG_TILE=3 python tools/kcycle_probe.py gpu 8192 --hermitian --restart 8 > gpurun_out/r3e_k8192_cpasync.txt 2>&1; tail -1 gpurun_out/r3e_k8192_cpasync.txt | cut -c1-400
QMG_TILE=3 QMG_HERM_MIN_NC=4 python tools/kcycle_probe.py gpu 8192 --hermitian --restart 8 > gpurun_out/r3e_k8192_cpasync_fine_stored.txt 2>&1; tail -1 gpurun_out/r3e_k8192_cpasync_fine_stored.txt | cut -c1-400
