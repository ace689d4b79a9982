#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs3.txt; }
: > gpurun_out/legs3.txt
leg t_pfx 240 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "prefix_shared" -n 4
leg prof_pairs 120 python tools/profile_pairs.py 20
tail -n 15 gpurun_out/t_pfx.log
cat gpurun_out/prof_pairs.log
cat gpurun_out/legs3.txt
