#!/bin/bash
for m in 0 1 0 1; do QMG_MR2=$m timeout 400 python tools/kcycle_probe.py gpu 4096 --hermitian --hermitian-setup --restart 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mr2=$m', {k: d.get(k) for k in ('iter', 'second_solve_s', 'check_relres', 'executed')})"; done
