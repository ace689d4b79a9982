"""Host-side profile of the AAT phase of one methyloxirane-shape step (where do the ~260 ms go?)."""
import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, apyib_b200
from apyib_b200 import aats
cfg = apyib_b200.config
cfg.VERBOSE = False
wl = bench.WORKLOADS[os.environ.get("WL", "methyloxirane")]
cfg.AAT_ALGORITHM = wl["algorithm"]
cfg.RETURN_DEVICE = True
work = bench.prepare(wl)
for _ in range(3):
    bench.gpu_step(work)
torch.cuda.synchronize()
print("graphs:", {k[7:10]: type(v).__name__ for k, v in aats._block_graphs.items()})
ph = []
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
bench.gpu_step(work, phases=ph)
pr.disable()
torch.cuda.synchronize()
print("step %.3f s" % (time.perf_counter() - t0), {b[0]: round(a[1].elapsed_time(b[1]), 1) for a, b in zip(ph[:-1], ph[1:])})
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
