#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for n in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/s_meth_n$n.log 2>&1
echo "n$n rc=$?"
tail -1 gpurun_out/s_meth_n$n.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'checksum', d['aat_checksum'])
print(' phases', d['phases_max_over_ranks']); print(' steps', d['step_times_s'])
" || tail -30 gpurun_out/s_meth_n$n.log
done
