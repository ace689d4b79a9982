"""Which switch makes the device-resident and the e2e leg of the methyloxirane step disagree?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import device as dev, aats
cfg = apyib_b200.config
cfg.VERBOSE = False
wl = bench.WORKLOADS[os.environ.get("WL", "methyloxirane")]
cfg.AAT_ALGORITHM = wl["algorithm"]
work = bench.prepare(wl)
dev.pin_host_inputs([work["w0"]] + list(work["pts"].values()))

def run(resident, conc, graph, stage):
    cfg.RETURN_DEVICE, cfg.SOLVE_CONCURRENT, cfg.AAT_USE_GRAPH = resident, conc, graph
    if not resident:
        bench.drop_device_caches(work)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    I = bench.gpu_step(work)
    torch.cuda.synchronize()
    return I, time.perf_counter() - t0

ref, _ = run(True, False, False, False)
ref2, _ = run(True, False, False, False)
print("baseline repeat diff", np.abs(ref - ref2).max(), "max|I|", np.abs(ref).max())
for resident in (True, False):
    for conc in (True,):
        for graph in (False, True):
            for stage in (False, True):
                for rep in range(4):
                    I, dt = run(resident, conc, graph, stage)
                    print("resident=%d conc=%d graph=%d stage=%d rep=%d: diff %.3e  %.3f s" % (resident, conc, graph, stage, rep, np.abs(I - ref).max(), dt), flush=True)
