#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs8.txt; }
: > gpurun_out/legs8.txt
leg t_all8 480 python -m pytest tests -m gpu -x -q -n 4
leg b_v9 240 python bench.py
leg b_mox 360 python bench.py --workload methyloxirane --steps 1 --warmup 1
leg ncu_list5 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v5.csv python bench.py --profile-step --warmup 1 --aat-graph 0
tail -n 6 gpurun_out/t_all8.log
tail -n 1 gpurun_out/b_v9.log | cut -c 1-300
tail -n 1 gpurun_out/b_mox.log | cut -c 1-900
tail -n 2 gpurun_out/ncu_list5.log
cat gpurun_out/legs8.txt
