#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r4b_bench1.json 2> gpurun_out/r4b_bench1.err; echo "rc $?"; tail -2 gpurun_out/r4b_bench1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 > gpurun_out/r4b_bench2.json 2> gpurun_out/r4b_bench2.err; echo "rc $?"; tail -2 gpurun_out/r4b_bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/r4b_bench1.json", "gpurun_out/r4b_bench2.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
    print(' shard_parity', d.get('shard_parity') and {k:d['shard_parity'][k] for k in ('max_slab_rel_error','ok_on_every_rank')}, d.get('shard_parity') and d['shard_parity']['slabs_result'])
    for k in ('kcycle','kcycle_strong'):
        if d.get(k): print(' ',k, {a:d[k].get(a) for a in ('iter','seconds','seconds_stored_blocks','setup_seconds','lattice','per_level_ops','per_level_ops_executed','check_relres','hbm_in_use_gb','link_compressed_levels')})
    if d.get('cpu_baseline'): print(' cpu', d['cpu_baseline'].get('kcycle_extrapolated'))
PY
