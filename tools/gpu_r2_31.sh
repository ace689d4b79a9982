#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg t_sub 900 python -m pytest tests/test_gpu_aat.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4
tail -3 gpurun_out/t_sub.log
leg b_meth 900 python bench.py --no-cpu-baseline
tail -1 gpurun_out/b_meth.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print(' phases', d['phases_max_over_ranks']); print(' e2e phases', d['e2e']['phases_max_over_ranks']); print(' steps', d['step_times_s'])
"
