#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs7.txt; }
: > gpurun_out/legs7.txt
leg t_solv 240 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4
leg t_all7 480 python -m pytest tests -m gpu -x -q -n 4
leg b_v8 240 python bench.py
leg hostprof8 150 python tools/host_profile.py --graph 1
tail -n 12 gpurun_out/t_solv.log
tail -n 6 gpurun_out/t_all7.log
tail -n 1 gpurun_out/b_v8.log | cut -c 1-300
head -12 gpurun_out/hostprof8.log
cat gpurun_out/legs7.txt
