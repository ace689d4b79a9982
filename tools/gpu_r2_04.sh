#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg t_all 900 python -m pytest tests -m gpu -x -q -n 4
tail -5 gpurun_out/t_all.log
leg b_small 200 python bench.py --workload small --steps 2 --warmup 1
tail -c 1500 gpurun_out/b_small.log
leg b_meth 900 python bench.py
tail -c 6000 gpurun_out/b_meth.log
leg b_h2o2 300 python bench.py --workload h2o2
tail -c 3000 gpurun_out/b_h2o2.log
