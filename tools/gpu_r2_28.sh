#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/d_meth.log 2>&1; echo "driver-like bench rc=$?"
tail -1 gpurun_out/d_meth.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
r=d['roofline']; print(' roof', r['kernel'], r['frac'], r['share_of_step'], r['traffic'], r['launches_of_signature'])
print(' steps', d['step_times_s'])
"
