#!/bin/bash
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 > gpurun_out/r3j_bench8.json 2> gpurun_out/r3j_bench8.err; echo "rc $?"; tail -5 gpurun_out/r3j_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3j_bench8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e')})
print(d['shard_parity'])
for k in ('kcycle','kcycle_strong'):
    print(k, {a:d[k].get(a) for a in ('iter','seconds','seconds_stored_blocks','setup_seconds','lattice','per_level_ops_executed','check_relres','precond_apply_s','hbm_in_use_gb')})
PY
