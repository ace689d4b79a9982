#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg t_conc 300 python -m pytest tests/test_gpu_solvers.py -m gpu -x -q -k "concurrent_identical or diis_drop_in"
tail -3 gpurun_out/t_conc.log
leg b_meth 900 python bench.py --no-cpu-baseline
leg b_mp2 600 python bench.py --no-cpu-baseline --method MP2
for f in b_meth b_mp2; do tail -1 gpurun_out/$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f', 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print(' phases', d['phases_max_over_ranks']); print(' e2e phases', d['e2e']['phases_max_over_ranks']); print(' steps', d['step_times_s'])
r=d['roofline']
print(' roof', r['kernel'], r['frac'], r['share_of_step'])
print(' top', r['top_signatures_ms'], r['eager_step_s'])
print(' aat top', r.get('top_signatures_aat_phase_ms'), r.get('timed_kernel_ms_aat_phase'), r.get('timed_kernel_ms_total'))
" || tail -20 gpurun_out/$f.log; done
