#!/bin/bash
# One gpurun call: new tests, bench (AAT graph replay off / on), full GPU suite, launch list, nso sweep.
# Every leg has its own timeout and log under gpurun_out/; later legs run even if earlier ones fail.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs.txt; }
: > gpurun_out/legs.txt
leg t_new 300 python -m pytest tests -m gpu -x -q -k "perturbed_cid or parallel_apts or hessian_pipeline or gradient_drivers or block_graph or apt_pipeline"
leg b0 240 python bench.py --aat-graph 0
leg b1 240 python bench.py --aat-graph 1
leg t_all 480 python -m pytest tests -m gpu -x -q -n 4
APYIB_B200_AAT_GRAPH=1 leg t_aat_graph 300 python -m pytest tests/test_gpu_aat.py -m gpu -x -q -n 4
leg ncu 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python bench.py --profile-step --warmup 2 --aat-graph 1
leg sweep 200 python tools/bench_contraction.py --big
tail -n 3 gpurun_out/t_new.log gpurun_out/t_all.log gpurun_out/t_aat_graph.log
tail -n 1 gpurun_out/b0.log gpurun_out/b1.log | cut -c 1-1500
cat gpurun_out/legs.txt
