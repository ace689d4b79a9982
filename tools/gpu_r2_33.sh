#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg t_sub 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_aat.py -m gpu -x -q -n 4
tail -3 gpurun_out/t_sub.log
python tools/profile_pairs2.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:det_pairs -s 3 -c 1 -f -o gpurun_out/r02_det_pairs_v3 python tools/profile_pairs2.py > gpurun_out/ncu_pairs.log 2>&1; echo "ncu rc=$?"
leg b_h2o2 300 python bench.py --workload h2o2 --no-cpu-baseline
tail -1 gpurun_out/b_h2o2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
r=d['roofline']; print(' roof', r['kernel'], r['frac'], r['share_of_step'], r.get('determinants_per_s'))
"
