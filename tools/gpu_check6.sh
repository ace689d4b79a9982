#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs6.txt; }
: > gpurun_out/legs6.txt
leg t_stack 240 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "stack_entry or prefix_shared" -n 4
leg t_all6 480 python -m pytest tests -m gpu -x -q -n 4
leg b_v7 240 python bench.py
leg hostprof7 150 python tools/host_profile.py --graph 1
tail -n 8 gpurun_out/t_stack.log
tail -n 6 gpurun_out/t_all6.log
tail -n 1 gpurun_out/b_v7.log | cut -c 1-600
head -12 gpurun_out/hostprof7.log
cat gpurun_out/legs6.txt
