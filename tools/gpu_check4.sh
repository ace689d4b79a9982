#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs4.txt; }
: > gpurun_out/legs4.txt
export APYIB_B200_LU_PREFIX=1
leg t_all_pfx 480 python -m pytest tests -m gpu -x -q -n 4
leg b_pfx 240 python bench.py
leg hostprof_pfx 150 python tools/host_profile.py --graph 1
tail -n 6 gpurun_out/t_all_pfx.log
tail -n 1 gpurun_out/b_pfx.log | cut -c 1-1200
head -24 gpurun_out/hostprof_pfx.log
cat gpurun_out/legs4.txt
