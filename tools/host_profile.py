"""Where the wall clock of one bench step goes on the HOST side: phase timings (each closed by a device
synchronise) and a cProfile of whole steps.  python tools/host_profile.py [--graph 0|1] [--workload h2o2]"""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import _lib
from apyib_b200.aats import AAT

cfg = apyib_b200.config
cfg.VERBOSE = False
cfg.AAT_USE_GRAPH = "--graph" in sys.argv and sys.argv[sys.argv.index("--graph") + 1] == "1"
wl = bench.WORKLOADS[sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "h2o2"]
work = bench.prepare(wl)
cfg.RETURN_DEVICE = True
for _ in range(3):
    bench.gpu_step(work)
torch.cuda.synchronize()


def clock(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n0 = _lib.LAUNCHES[0]
    r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-34s host %8.2f ms   +drain %7.2f ms   launches %5d" % (label, 1e3 * (t1 - t0), 1e3 * (t2 - t1), _lib.LAUNCHES[0] - n0))
    return r


par, w0, natom = work["par"], work["w0"], work["natom"]
n3 = 3 * natom
from apyib_b200.ci_wfn import solve_many, ci_wfn, solve_batch
pts = list(work["pts"])
for rep in range(2):
    print("--- repetition %d (graph replay of AAT stacks: %s)" % (rep, cfg.AAT_USE_GRAPH))
    cis = clock("ci_wfn ctors (F_MO, ERI_MO) x%d" % (len(pts) + 1), lambda: [ci_wfn(par, w) for w in [w0] + [work["pts"][p] for p in pts]])
    del cis
    sols = clock("solve_many (ctors + batched CISD)", lambda: solve_many("CISD", par, [w0] + [work["pts"][p] for p in pts]))
    T0 = [1, sols[0][1], sols[0][2]]
    mine = {p: [1, r[1], r[2]] for p, r in zip(pts, sols[1:])}
    T = lambda k, i, s: mine[(k, i, s)]
    W = lambda k, i, s: work["pts"][(k, i, s)]
    A = clock("AAT ctor (175 MO overlaps)", lambda: AAT(
        par, w0, w0.C, w0.H.basis_set, T0,
        [W("R", a, +1).C for a in range(n3)], [W("R", a, -1).C for a in range(n3)],
        [W("R", a, +1).H.basis_set for a in range(n3)], [W("R", a, -1).H.basis_set for a in range(n3)],
        [T("R", a, +1) for a in range(n3)], [T("R", a, -1) for a in range(n3)],
        [W("B", b, +1).C for b in range(3)], [W("B", b, -1).C for b in range(3)],
        [W("B", b, +1).H.basis_set for b in range(3)], [W("B", b, -1).H.basis_set for b in range(3)],
        [T("B", b, +1) for b in range(3)], [T("B", b, -1) for b in range(3)], bench.H_R, bench.H_B))
    clock("spatial norms", lambda: A._spatial_norms("full"))
    clock("spatial amplitude sets", lambda: A._spatial_amps("full"))
    clock("first element (uu, up/un, pu/nu, blk4)", lambda: A.compute_spatial_aats(0, 0))
    clock("elements (0,1), (0,2)", lambda: [A.compute_spatial_aats(0, b) for b in (1, 2)])
    clock("remaining %d elements" % (3 * n3 - 3), lambda: [A.compute_spatial_aats(a, b) for a in range(1, n3) for b in range(3)])
    clock("whole gpu_step", lambda: bench.gpu_step(work))

pr = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.perf_counter()
pr.enable()
for _ in range(3):
    bench.gpu_step(work)
torch.cuda.synchronize()
pr.disable()
print("3 profiled steps: %.3f s (cProfile overhead included)" % (time.perf_counter() - t0))
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()))
