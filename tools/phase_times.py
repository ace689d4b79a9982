"""Wall-clock / GPU-time breakdown of one bench step (solves vs AAT assembly)."""
import os, sys, time, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import _lib
apyib_b200.config.VERBOSE = False
apyib_b200.config.AAT_ALGORITHM = os.environ.get("ALGO", "lu")
wl = bench.WORKLOADS[os.environ.get("WL", "h2o2")]
work = bench.prepare(wl)
apyib_b200.config.RETURN_DEVICE = True
bench.gpu_step(work)            # warm-up
torch.cuda.synchronize()
par = work["par"]
from apyib_b200.ci_wfn import solve_many
t0 = time.perf_counter()
sols = solve_many("CISD", par, [work["w0"]] + list(work["pts"].values()))
torch.cuda.synchronize()
print("solve_many (all %d points): %.3f s" % (len(sols), time.perf_counter() - t0))
del sols
for graph in ((True, False) if os.environ.get("WL", "h2o2") != "methyloxirane" else ()):
    apyib_b200.config.USE_CUDA_GRAPH = graph
    t0 = time.perf_counter(); n0 = _lib.LAUNCHES[0]
    its = []
    for w in [work["w0"]] + list(work["pts"].values()):
        c = apyib_b200.ci_wfn(par, w)
        t1 = time.perf_counter()
        c.solve_CISD(); its.append(c.iterations)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("graph=%s: %d solves %.3f s (%.1f ms each, %d launches, iterations %s)" % (graph, len(its), dt, 1e3 * dt / len(its), _lib.LAUNCHES[0] - n0, sorted(set(its))))
apyib_b200.config.USE_CUDA_GRAPH = True
w = work["w0"]
t0 = time.perf_counter(); c = apyib_b200.ci_wfn(par, w); torch.cuda.synchronize(); t1 = time.perf_counter()
c.solve_CISD(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("ctor (F_MO + ERI_MO) %.2f ms, solve %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
apyib_b200.config.USE_CUDA_GRAPH = False
import cProfile, pstats
apyib_b200.config.TIMING = {}
t0 = time.perf_counter(); n0 = _lib.LAUNCHES[0]
pr = cProfile.Profile(); pr.enable()
I = bench.gpu_step(work)
pr.disable()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tm = apyib_b200.config.TIMING; apyib_b200.config.TIMING = None
agg = collections.Counter(); cnt = collections.Counter()
for k, ev in tm.items():
    key = k.split("[")[0] + ("[" + k.split("[")[1].split(" ")[0] if k.startswith("contract") else "")
    agg[k if k.startswith("det") else key] += sum(a.elapsed_time(b) for a, b in ev); cnt[k if k.startswith("det") else key] += len(ev)
print("step with timing hooks: %.3f s, launches %d" % (dt, _lib.LAUNCHES[0] - n0))
for k, v in agg.most_common(12):
    print("  %-40s n=%6d %9.1f ms" % (k, cnt[k], v))

import re
print("top contraction signatures (eager step):")
rows = []
for k, ev in tm.items():
    m = re.match(r"contract(_tma)?\[(f64|c128) (\d+)x(\d+)x(\d+) b(\d+)\]", k)
    if not m:
        continue
    ms = sum(a.elapsed_time(b) for a, b in ev)
    M, N, K, nb = (int(m.group(i)) for i in (3, 4, 5, 6))
    fl = (8.0 if m.group(2) == "c128" else 2.0) * M * N * K * nb * len(ev)
    rows.append((ms, k, len(ev), fl / (ms * 1e-3) / 1e12))
for ms, k, n, tf in sorted(rows, reverse=True)[:24]:
    print("  %-46s n=%5d %8.1f ms  %6.2f TFLOP/s" % (k, n, ms, tf))
print("  total contraction time %.1f ms" % sum(r[0] for r in rows))

pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
