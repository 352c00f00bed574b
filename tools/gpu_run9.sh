#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --kcycle-L 0 --no-cpu > gpurun_out/r3i_bench_stencil.json 2> gpurun_out/r3i_bench_stencil.err; echo "rc $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3i_bench_stencil.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])"
python tools/kernel_probe.py --only stencil --reps 10 > gpurun_out/r3i_stencil_probe.log 2>&1; cat gpurun_out/r3i_stencil_probe.log
python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r3i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stencil_kernel -s 3 -c 2 -o gpurun_out/r3i_stencil python bench.py --steps 3 --warmup 3 --kcycle-L 0 --no-cpu > gpurun_out/r3i_stencil_ncu.log 2>&1
python tools/setup_profile.py 1024 > /dev/null 2>&1 &&
ncu --set full --clock-control none -k regex:"coarse_build_kernel|block_ortho_kernel" -c 4 -o gpurun_out/r3i_setup_kernels python tools/setup_profile.py 1024 > gpurun_out/r3i_setup_ncu.log 2>&1
python tools/kernel_probe.py --only transfer --reps 3 > /dev/null 2>&1 &&
ncu --set full --clock-control none -k regex:restrict_kernel -s 3 -c 1 -o gpurun_out/r3i_restrict python tools/kernel_probe.py --only transfer --reps 3 > gpurun_out/r3i_restrict_ncu.log 2>&1
