"""One launch of the prefix-shared LU kernel at the H2O2/6-31G shape (n = 9, 2808 x 2808 doubles x doubles table, 60 overlaps,
one amplitude vector) -- target for `ncu --set full -k regex:det_pairs`."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apyib_b200
from apyib_b200._lib import lib, check
from apyib_b200.aats import _Tables
from apyib_b200.device import to_device, empty, ptr, stream_ptr
n, nv, nS = 9, 13, int(os.environ.get("NS", 60))
ns = n + nv
rng = np.random.default_rng(0)
S = to_device(np.stack([np.eye(ns) + 1e-4 * (rng.standard_normal((ns, ns)) + 0.1j * rng.standard_normal((ns, ns))) for _ in range(nS)]), torch.complex128)
T = _Tables.get(n, 0, nv)
rows = T.L[2]; cs, sg, ix = T.LS[2]; gl, cand, nc = T.PFX[2]
nrow, ncol = rows.shape[0], cs.shape[0]
Y = to_device(rng.standard_normal((1, ncol)) + 1j * rng.standard_normal((1, ncol)), torch.complex128)
Z = empty((nS, 1, nrow), torch.complex128)
work = empty((nS * int(lib.apyib_det_matvec_pairs_work_len(nrow, ncol // gl, 1, n, 2, ns, nc)),), torch.complex128)
def run():
    check(lib.apyib_det_matvec_pairs_stack(ptr(S), nS, ns, n, 2, ptr(rows), nrow, ptr(cs), ptr(sg), ptr(ix), ncol, gl, ptr(cand), nc,
                                           ptr(Y), 0, 1, ptr(Z), ptr(work), stream_ptr()))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("det_pairs n=9 k=2 %dx%d nS=%d: %.3f ms per launch set, %.3g det/s" % (nrow, ncol, nS, ms, nS * nrow * ncol / ms * 1e3))
