#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs5.txt; }
: > gpurun_out/legs5.txt
leg b_v6 240 python bench.py
leg ncu_full 150 env NCU=1 ncu --set full --clock-control none --import-source on -k regex:det_pairs --launch-count 3 -o gpurun_out/det_pairs_v1 -f python tools/profile_pairs.py
leg ncu_list 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv python bench.py --profile-step --warmup 1 --aat-graph 0
tail -n 1 gpurun_out/b_v6.log | cut -c 1-3000
tail -3 gpurun_out/ncu_full.log; tail -3 gpurun_out/ncu_list.log
cat gpurun_out/legs5.txt
