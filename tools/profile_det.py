"""Runs the fused LU determinant kernel on the bench shape (n=9, 2808 x 2808 substituted
matrices of a 22 x 22 overlap) a few times -- target for `ncu -k regex:det_kernel`."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apyib_b200.aats import _Tables, _det_matvec
from apyib_b200.device import to_device

no, nf, nv = int(os.environ.get("NO", 9)), 0, int(os.environ.get("NV", 13))
nbf = no + nv
rng = np.random.default_rng(0)
S = to_device(np.eye(nbf) + 1e-4 * (rng.standard_normal((nbf, nbf)) + 0.1j * rng.standard_normal((nbf, nbf))), torch.complex128)
T = _Tables.get(no, nf, nv)
P = T.n2
Y = to_device(rng.standard_normal((1, P)) + 1j * rng.standard_normal((1, P)), torch.complex128)
from apyib_b200._lib import lib, check
kernels = [int(k) for k in os.environ.get("KERNELS", "0,1,2").split(",")]     # 0 = thread-per-matrix, 1 = sub-warp
ref = None
import apyib_b200
for which in kernels:            # 0 = thread-per-matrix, 1 = sub-warp, 2 = thread-per-matrix + factorisation reuse
    check(lib.apyib_det_set_kernel(which & 1))
    apyib_b200.config.LU_REUSE = which == 2
    _dm = _det_matvec
    if which == 2:
        _det_matvec = lambda S_, n_, r_, c_, Y_: _dm(S_, n_, r_, c_, Y_, T.LS[2])
    for _ in range(3):
        Z = _det_matvec(S, no, T.L[2], T.L[2], Y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        Z = _det_matvec(S, no, T.L[2], T.L[2], Y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if ref is None:
        ref = Z.clone()
    _det_matvec = _dm
    print("kernel=%d n=%d P=%d  %.3f ms/launch  %.3e dets/s  %.2f TFLOP/s (8/3 n^3)  maxdiff vs first %.2e (scale %.2e)"
          % (which, no, P, ms, P * P / ms * 1e3, P * P * 8 / 3 * no ** 3 / ms * 1e3 / 1e12,
             float((Z - ref).abs().max()), float(ref.abs().max())))
