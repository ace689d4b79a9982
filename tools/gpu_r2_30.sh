#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { name=$1; t=$2; shift 2; s=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(( $(date +%s) - s ))s"; }
leg t_sub 900 python -m pytest tests/test_gpu_aat.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4
tail -3 gpurun_out/t_sub.log
timeout 300 python tools/profile_aat2.py 2>&1 | tail -5
python - <<'PY'
import cProfile, pstats, os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench, apyib_b200
from apyib_b200.aats import AAT
from apyib_b200.energy import correlated_many
cfg = apyib_b200.config; cfg.VERBOSE = False; cfg.AAT_ALGORITHM = "factorized"; cfg.RETURN_DEVICE = True
wl = bench.WORKLOADS["methyloxirane"]; work = bench.prepare(wl)
par, w0, natom = work["par"], work["w0"], work["natom"]; n3 = 3 * natom
from apyib_b200.utils import ao_prefetch
ao_prefetch([w0] + list(work["pts"].values())); torch.cuda.synchronize()
W = lambda k, i, s: work["pts"][(k, i, s)]
T0 = [1, 0, 0]
def mk():
    return AAT(par, w0, w0.C, w0.H.basis_set, T0,
            [W("R", a, +1).C for a in range(n3)], [W("R", a, -1).C for a in range(n3)],
            [W("R", a, +1).H.basis_set for a in range(n3)], [W("R", a, -1).H.basis_set for a in range(n3)],
            [T0] * n3, [T0] * n3, [W("B", b, +1).C for b in range(3)], [W("B", b, -1).C for b in range(3)],
            [W("B", b, +1).H.basis_set for b in range(3)], [W("B", b, -1).H.basis_set for b in range(3)],
            [T0] * 3, [T0] * 3, 1e-4, 1e-4)
mk(); mk(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); mk(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
PY
