#!/bin/bash
# round 2, call 2: phase breakdown + launch list of the methyloxirane-shape step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
WL=methyloxirane ALGO=factorized timeout 600 python tools/phase_times.py > gpurun_out/phase_meth.log 2>&1
echo "phase rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_meth.csv python bench.py --workload methyloxirane --profile-step --warmup 1 > gpurun_out/ncu_meth.log 2>&1
echo "ncu rc=$?"
tail -60 gpurun_out/phase_meth.log
