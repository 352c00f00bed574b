#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r4a_pytest.log 2>&1; tail -6 gpurun_out/r4a_pytest.log
python tools/kcycle_probe.py gpu 4096 8192 --hermitian --restart 8 --profile > gpurun_out/r4a_k.txt 2>&1; grep -E "PROFILE\] (qmg_st|qmg_gcr|qmg_kry|qmg_pro|qmg_res|qmg_mul|total)|second_solve" gpurun_out/r4a_k.txt | cut -c1-330
