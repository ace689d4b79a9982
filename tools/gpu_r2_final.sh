#!/bin/bash
# what the driver runs at round end, in the same form: single-process GPU suite, smoke, bench (both arms)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
s=$(date +%s); timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s) - s ))s"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
s=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.log 2>&1; echo "bench rc=$? $(( $(date +%s) - s ))s"
tail -1 gpurun_out/final_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
r=d['roofline']; print(' roof', r['kernel'], r['frac'], r['share_of_step'], r['traffic'])
print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['sample_wall_s'])
"
