"""Concurrent-mode deviation of the first point of a group: are its MO integrals or its iterations off?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
cw = sys.modules["apyib_b200.ci_wfn"]
cfg = apyib_b200.config
cfg.VERBOSE = False
wl = bench.WORKLOADS["methyloxirane"]
work = bench.prepare(wl)
par = work["par"]
wfns = [work["w0"]] + list(work["pts"].values())
cfg.RETURN_DEVICE = True
cw._DEBUG_KEEP = []

def run(conc):
    cfg.SOLVE_CONCURRENT = conc
    cw._DEBUG_KEEP.clear()
    torch.cuda.synchronize()
    res = cw.solve_many("CISD", par, wfns)
    torch.cuda.synchronize()
    cis = cw._DEBUG_KEEP[0]
    ints = [(c._F_dev.clone(), c._ERI_dev.clone() if k < 8 else c._ERI_dev[:1, :1].clone(), c.iterations) for k, c in enumerate(cis)]
    cw._DEBUG_KEEP.clear()
    return res, ints

cfg.USE_TMA = os.environ.get('TMA', '1') == '1'
print('USE_TMA', cfg.USE_TMA)
ref, ref_i = run(False)
for rep in range(6):
    res, ints = run(True)
    bad = []
    for k, (a, b) in enumerate(zip(ref, res)):
        d = max(float((a[1] - b[1]).abs().max()), float((a[2] - b[2]).abs().max()))
        dF = float((ref_i[k][0] - ints[k][0]).abs().max()); dE = float((ref_i[k][1] - ints[k][1]).abs().max())
        if d > 1e-12 or dF > 0 or dE > 0:
            bad.append((k, "dt=%.2e dF=%.2e dERI=%.2e its %d/%d" % (d, dF, dE, ref_i[k][2], ints[k][2])))
    print("rep %d:" % rep, bad[:5], flush=True)
