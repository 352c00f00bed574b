#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/baseline_configs_probe.py > gpurun_out/r5z_baseline_configs.txt 2>&1; cat gpurun_out/r5z_baseline_configs.txt | cut -c1-500
