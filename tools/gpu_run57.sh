#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py tests/test_shard_gpu.py -m gpu -q -x -k "fused or kcycle or n22 or staggered or n16 or loopback" 2>&1 | tail -6
timeout 400 python tools/kcycle_probe.py gpu 8192 --hermitian --hermitian-setup --restart 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mr2', {k: d.get(k) for k in ('iter', 'second_solve_s', 'second_solve_iter', 'setup_seconds', 'check_relres', 'executed')})"
