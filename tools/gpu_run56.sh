#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_host_gpu.py -m gpu -q -x -k "packed or transfer or kcycle or n22 or staggered or n16" 2>&1 | tail -8
for m in 0 1; do QMG_PACKED_TRANSFER=$m timeout 400 python tools/kcycle_probe.py gpu 8192 --hermitian --hermitian-setup --restart 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('packed $m', {k: d.get(k) for k in ('iter', 'second_solve_s', 'second_solve_iter', 'setup_seconds', 'check_relres', 'executed')})"; done
