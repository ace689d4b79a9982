#!/bin/bash
# N=2: sharded bench at the target shape (NCCL)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/b_meth_n2.log 2>&1
echo "n2 rc=$?"
tail -1 gpurun_out/b_meth_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print(' phases', d['phases_max_over_ranks']); print(' e2e phases', d['e2e']['phases_max_over_ranks']); print(' steps', d['step_times_s'])
" || tail -30 gpurun_out/b_meth_n2.log
