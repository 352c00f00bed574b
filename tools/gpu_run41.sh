#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bicgstab_fused" 2>&1 | tail -6
for r in 8 12; do timeout 400 python tools/kcycle_probe.py gpu 8192 --hermitian --hermitian-setup --restart $r 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('restart $r', {k: d.get(k) for k in ('iter', 'seconds', 'second_solve_s', 'second_solve_iter', 'setup_seconds', 'check_relres', 'executed')})"; done
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
