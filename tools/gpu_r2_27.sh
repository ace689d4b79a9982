#!/bin/bash
# round-2 artefacts: final bench lines (CISD / MP2 / H2O2 / sweep / reference arm), ncu captures and launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg f_meth 600 python bench.py
leg f_h2o2 300 python bench.py --workload h2o2
leg f_mp2 300 python bench.py --method MP2
leg f_sweep 600 python bench.py --workload sweep --steps 5
leg f_ref 600 python bench.py --impl reference --steps 3 --warmup 1
cap() { name=$1; shift; timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract -s 2 -c 1 -f -o gpurun_out/r02b_$name env "$@" python tools/profile_ring.py > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"; }
cap ring_f64 NB=31 ONLY=skbjc,sikca NOLADDER=1
cap skinny_f64 NB=31 ONLY=sjabc,sic NOLADDER=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_meth_final.csv python bench.py --profile-step --warmup 2 > gpurun_out/ncu_meth_final.log 2>&1
echo "launch list rc=$?"
for f in f_meth f_h2o2 f_mp2 f_sweep f_ref; do echo "== $f"; tail -1 gpurun_out/$f.log | cut -c1-600; done
