"""Where does the e2e leg lose time against the device-resident leg?  (methyloxirane shape)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import device as dev, utils
cw = sys.modules["apyib_b200.ci_wfn"]
cfg = apyib_b200.config
cfg.VERBOSE = False
wl = bench.WORKLOADS["methyloxirane"]
work = bench.prepare(wl)
par = work["par"]
wfns = [work["w0"]] + list(work["pts"].values())
print("pinned:", dev.pin_host_inputs(wfns))
x = torch.from_numpy(np.asarray(wfns[3].H.ERI)); print("ERI is_pinned:", x.is_pinned(), x.dtype, x.numel() * 8 / 1e6, "MB")
T0 = [0.0]
def stamp(msg):
    print("  %-46s %7.3f s" % (msg, time.perf_counter() - T0[0]), flush=True)
# pure upload time
for rep in range(2):
    bench.drop_device_caches(work); torch.cuda.synchronize()
    T0[0] = time.perf_counter()
    utils.ao_prefetch(wfns); stamp("prefetch enqueued (%d points)" % len(wfns))
    dev.copy_stream().synchronize(); stamp("uploads complete")
# instrumented e2e solve_many: wrap _run_jobs' sequential loop
orig_collect = cw._collect
def collect(eng, E, singles):
    torch.cuda.synchronize(); stamp("job done (nb=%d, %s) iterations=%s" % (eng.nb, eng.dtype, max(eng.iterations)))
    r = orig_collect(eng, E, singles); stamp("  collected"); return r
cw._collect = collect
for resident in (True, False, False):
    cfg.RETURN_DEVICE = resident
    if not resident:
        bench.drop_device_caches(work)
    torch.cuda.synchronize()
    print("resident =", resident)
    T0[0] = time.perf_counter()
    res = cw.solve_many("CISD", par, wfns)
    torch.cuda.synchronize(); stamp("solve_many done")
    del res
