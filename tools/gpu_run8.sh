#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r3h_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3h_bench_ref.json 2> gpurun_out/r3h_bench_ref.err; echo "ref rc $?"
python tools/setup_profile.py 4096 > gpurun_out/r3h_setup4096.txt 2>&1; grep -v gpurun gpurun_out/r3h_setup4096.txt | head -30
python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r3h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stencil_kernel -s 3 -c 2 -o gpurun_out/r3h_stencil python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r3h_stencil_ncu.log 2>&1
python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r3h_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3h_launches.csv python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r3h_launches_ncu.log 2>&1
