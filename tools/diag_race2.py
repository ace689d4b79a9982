"""Localise the concurrent-mode race: compare the amplitudes of solve_many (not the AAT tensor) against a sequential run."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import device as dev
from apyib_b200.ci_wfn import solve_many
cfg = apyib_b200.config
cfg.VERBOSE = False
wl = bench.WORKLOADS[os.environ.get("WL", "methyloxirane")]
work = bench.prepare(wl)
par = work["par"]
wfns = [work["w0"]] + list(work["pts"].values())
cfg.RETURN_DEVICE = True

def run(conc, graph, chunk, stage):
    cfg.SOLVE_CONCURRENT, cfg.USE_CUDA_GRAPH, cfg.SOLVE_CHUNK = conc, graph, chunk
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solve_many("CISD", par, wfns)
    torch.cuda.synchronize()
    return res, time.perf_counter() - t0

ref, _ = run(False, True, 32, False)
def diff(res):
    bad = []
    for k, (a, b) in enumerate(zip(ref, res)):
        d = max(float((a[1] - b[1]).abs().max()), float((a[2] - b[2]).abs().max()), abs(complex(a[0]) - complex(b[0])))
        if d > 0:
            bad.append((k, d))
    return bad
print("repeat sequential:", diff(run(False, True, 32, False)[0]))
for conc, graph, chunk, stage in ((True, True, 32, False), (True, False, 32, False), (True, True, 1000, False), (True, False, 1000, False),
                                  (True, True, 32, True), (True, False, 32, True), (True, True, 1000, True), (False, True, 32, True)):
    for rep in range(4):
        res, dt = run(conc, graph, chunk, stage)
        bad = diff(res)
        print("conc=%d solvergraph=%d chunk=%d stage=%d rep=%d: %d bad points %s  %.3f s" % (conc, graph, chunk, stage, rep, len(bad), bad[:6], dt), flush=True)
