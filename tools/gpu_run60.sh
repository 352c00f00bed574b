#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wilson_mf_tile -s 4 -c 1 -o gpurun_out/r6m_mf python tools/mf_probe.py 8192 > gpurun_out/r6m_ncu.log 2>&1
ls -la gpurun_out | tail -3
