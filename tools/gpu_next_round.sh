#!/bin/bash
# First GPU call of the next round: validate and measure the experimental paths that were written after the
# round-1 GPU budget was spent (all OFF by default):
#   APYIB_B200_PAIRS_NY1=1          single-vector specialisation of the prefix-shared LU kernel
#   APYIB_B200_SOLVE_CONCURRENT=1   real and complex batches of a molecule solved concurrently on two streams
#   APYIB_B200_SCF_DEVICE_JK=1      (2J - K)[D] of the host SCF on the device (untimed host part; test only)
# gpurun --timeout 600 -- 'bash tools/gpu_next_round.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/legs_next.txt; }
: > gpurun_out/legs_next.txt
APYIB_B200_EXPERIMENTAL=1 leg t_exp 240 python -m pytest tests -m gpu -x -q -k "single_vector or concurrent_dtype or device_jk"
leg b_base 200 python bench.py
APYIB_B200_PAIRS_NY1=1 leg b_ny1 200 python bench.py
APYIB_B200_SOLVE_CONCURRENT=1 leg b_conc 200 python bench.py
APYIB_B200_PAIRS_NY1=1 APYIB_B200_SOLVE_CONCURRENT=1 leg b_both 200 python bench.py
APYIB_B200_PAIRS_NY1=1 APYIB_B200_SOLVE_CONCURRENT=1 leg t_all_exp 400 python -m pytest tests -m gpu -x -q -n 4
for f in b_base b_ny1 b_conc b_both; do echo "$f: $(tail -n 1 gpurun_out/$f.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"], d["roofline"]["avg_ms"], d["step_times_s"])' 2>&1)"; done
tail -n 3 gpurun_out/t_exp.log gpurun_out/t_all_exp.log
cat gpurun_out/legs_next.txt
