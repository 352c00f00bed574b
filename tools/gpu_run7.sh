#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --L 2048 --kcycle-L 1024 --steps 5 --cpu-kcycle-L 64 > gpurun_out/r3g_bench1_small.json 2> gpurun_out/r3g_bench1_small.err; echo "rc $?"; tail -3 gpurun_out/r3g_bench1_small.err; cut -c1-1500 gpurun_out/r3g_bench1_small.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --L 2048 --kcycle-L 2048 --steps 5 > gpurun_out/r3g_bench2_small.json 2> gpurun_out/r3g_bench2_small.err; echo "rc $?"; tail -5 gpurun_out/r3g_bench2_small.err; cat gpurun_out/r3g_bench2_small.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','e2e','shard_parity')},indent=1))
for k in ('kcycle','kcycle_strong'):
    print(k, {a:d[k][a] for a in ('iter','seconds','setup_seconds','lattice','per_level_ops','per_level_ops_executed','check_relres')})
"
python -m pytest tests/test_shard_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "two_rank or tile_kernel_flavours" 2>&1 | tail -3
