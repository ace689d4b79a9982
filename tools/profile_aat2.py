"""Synchronised wall-clock breakdown of the AAT phase at the methyloxirane shape."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import aats
from apyib_b200.aats import AAT
from apyib_b200.energy import correlated_many
cfg = apyib_b200.config
cfg.VERBOSE = False
wl = bench.WORKLOADS[os.environ.get("WL", "methyloxirane")]
cfg.AAT_ALGORITHM = wl["algorithm"]
cfg.RETURN_DEVICE = True
work = bench.prepare(wl)
par, w0, natom = work["par"], work["w0"], work["natom"]
n3 = 3 * natom
allp = [("U", 0, 0)] + bench.aat_points(natom)
wf = lambda p: w0 if p[0] == "U" else work["pts"][p]
sols = correlated_many(par, [wf(p) for p in allp])
mine = {p: T for p, (_, T) in zip(allp, sols)}
T = lambda k, i, s: mine[(k, i, s)]
W = lambda k, i, s: work["pts"][(k, i, s)]
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t0 = sync()
    A = AAT(par, w0, w0.C, w0.H.basis_set, mine[("U", 0, 0)],
            [W("R", a, +1).C for a in range(n3)], [W("R", a, -1).C for a in range(n3)],
            [W("R", a, +1).H.basis_set for a in range(n3)], [W("R", a, -1).H.basis_set for a in range(n3)],
            [T("R", a, +1) for a in range(n3)], [T("R", a, -1) for a in range(n3)],
            [W("B", b, +1).C for b in range(3)], [W("B", b, -1).C for b in range(3)],
            [W("B", b, +1).H.basis_set for b in range(3)], [W("B", b, -1).H.basis_set for b in range(3)],
            [T("B", b, +1) for b in range(3)], [T("B", b, -1) for b in range(3)], bench.H_R, bench.H_B)
    t1 = sync()
    A._spatial_norms("full"); t2 = sync()
    A._spatial_amps("full"); t3 = sync()
    A.compute_spatial_aats(0, 0); t4 = sync()      # uu, up/un (6), pu/nu (60, 2 chunks), blk4(alpha=0)
    for a in range(1, n3):
        A.compute_spatial_aats(a, 0)
    t5 = sync()
    for a in range(n3):
        for b in (1, 2):
            A.compute_spatial_aats(a, b)
    t6 = sync()
    print("rep %d: ctor(overlaps) %.1f  norms %.1f  amps %.1f  first element (uu+up/un+pu/nu+blk4) %.1f  29 blk4 stacks %.1f  rest %.1f  total %.1f ms"
          % (rep, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(t6-t5), 1e3*(t6-t0)), flush=True)
# finer: time each family in a fresh object
A._cache = {k: v for k, v in A._cache.items() if k[0] in ("norms", "amps")}
amps = A._spatial_amps("full"); A1, A2 = amps[1], amps[2]
for name in (("up", 0), ("pu", 0)):
    t0 = sync(); A._fill_family(name, "full", A1, A2); t1 = sync()
    print("family %s: %.1f ms" % (name[0], 1e3 * (t1 - t0)))
t0 = sync(); A._block(A.overlap_uu, A1["dR"], A2["dR"], A1["dH"], A2["dH"]); t1 = sync(); print("uu: %.1f ms" % (1e3*(t1-t0)))
