#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/s2_meth_n$n.log 2>&1
echo "n$n rc=$?"
tail -1 gpurun_out/s2_meth_n$n.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'], 'checksum', d['aat_checksum'])
print(' steps', d['step_times_s'])
" || tail -30 gpurun_out/s2_meth_n$n.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 2 --warmup 1 --impl reference 2>&1 | tail -1 | cut -c1-300
