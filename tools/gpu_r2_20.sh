#!/bin/bash
# ncu --set full captures of the dominant contraction kernels at the methyloxirane shape (one launch each)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cap() { name=$1; shift; timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract -s 2 -c 1 -f -o gpurun_out/r02_$name env "$@" python tools/profile_ring.py > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"; }
cap ring_f64 NB=31 ONLY=skbjc,sikca NOLADDER=1
cap skinny_f64 NB=31 ONLY=sjabc,sic NOLADDER=1
cap ladder_f64 NB=31 ONLY=spcd
cap ladder_c128 NB=6 CPLX=1 ONLY=spcd
NB=31 python tools/profile_ring.py > gpurun_out/ring_times_f64.log 2>&1
NB=6 CPLX=1 python tools/profile_ring.py > gpurun_out/ring_times_c128.log 2>&1
cat gpurun_out/ring_times_f64.log gpurun_out/ring_times_c128.log
ls -la gpurun_out/*.ncu-rep
