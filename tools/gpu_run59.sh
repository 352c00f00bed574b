#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r6l_pytest.log 2>&1; tail -4 gpurun_out/r6l_pytest.log
python bench.py > gpurun_out/r6l_bench1.json 2> gpurun_out/r6l_bench1.err; echo "rc $?"; tail -2 gpurun_out/r6l_bench1.err
python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r6l_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r6l_launches.csv python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r6l_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stencil_kernel|wilson_mf_tile" -s 3 -c 2 -o gpurun_out/r6l_stencil python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r6l_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
