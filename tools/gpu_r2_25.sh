#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg t_all 900 python -m pytest tests -m gpu -x -q -n 4
tail -4 gpurun_out/t_all.log
timeout 300 python tools/diag_race3.py 2>&1 | tail -7
leg b_h2o2 300 python bench.py --no-cpu-baseline --workload h2o2
leg b_meth 900 python bench.py --no-cpu-baseline
for f in b_meth b_h2o2; do tail -1 gpurun_out/$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f', 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print(' phases', d['phases_max_over_ranks']); print(' e2e phases', d['e2e']['phases_max_over_ranks']); print(' steps', d['step_times_s'])
" || tail -20 gpurun_out/$f.log; done
