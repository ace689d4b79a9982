"""DD table x vector at the H2O2/6-31G shape: per-matrix LU (with / without reuse) vs prefix-shared LU.
python tools/profile_pairs.py [reps]   (ncu target: set NCU=1 to run each variant once)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apyib_b200
from apyib_b200.aats import _Tables, _det_matvec
from apyib_b200.device import to_device, to_host
cfg = apyib_b200.config
no, nv = (9, 13) if len(sys.argv) < 4 else (int(sys.argv[2]), int(sys.argv[3]))
reps = 1 if os.environ.get("NCU") else (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
rng = np.random.default_rng(1)
ns = no + nv
S = to_device(np.eye(ns) + 1e-4 * (rng.standard_normal((ns, ns)) + 0.1j * rng.standard_normal((ns, ns))), torch.complex128)
T = _Tables.get(no, 0, nv)
res = {}
for name, rk, ck, ny in (("DD 2x2", 2, 2, 1), ("D x S", 2, 1, 2), ("S x D", 1, 2, 1)):
    rows, cols = T.L[rk], T.L[ck]
    Y = to_device(rng.standard_normal((ny, cols.shape[0])) + 1j * rng.standard_normal((ny, cols.shape[0])), torch.complex128)
    for label, reuse, prefix in (("plain", False, False), ("reuse", True, False), ("prefix", True, True)):
        cfg.LU_REUSE, cfg.LU_PREFIX = reuse, prefix
        Z = _det_matvec(S, no, rows, cols, Y, T.LS[ck], T.PFX[ck], ck)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            Z = _det_matvec(S, no, rows, cols, Y, T.LS[ck], T.PFX[ck], ck)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nd = rows.shape[0] * cols.shape[0]
        res[(name, label)] = to_host(Z)
        print("%-8s %-7s %9.4f ms  %10.3e det/s  algorithmic %7.2f TFLOP/s" % (name, label, ms, nd / ms * 1e3, nd * (8 / 3) * no ** 3 / ms / 1e9))
    ref = res[(name, "plain")]
    print("   max rel diff prefix vs plain: %.2e, reuse vs plain: %.2e" % (np.abs(res[(name, "prefix")] - ref).max() / np.abs(ref).max(),
                                                                        np.abs(res[(name, "reuse")] - ref).max() / np.abs(ref).max()))
