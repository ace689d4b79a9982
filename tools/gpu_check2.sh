#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs2.txt; }
: > gpurun_out/legs2.txt
leg t_full 200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q
leg hostprof1 150 python tools/host_profile.py --graph 1
leg hostprof0 120 python tools/host_profile.py --graph 0
tail -n 4 gpurun_out/t_full.log
cat gpurun_out/legs2.txt
