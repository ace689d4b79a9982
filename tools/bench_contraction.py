"""Roofline of the DMMA contraction kernel on the shapes of BASELINE config 3-5
(ladder O^2 x V^2 x V^2, ring (OV)^3, plain GEMM), complex128 and float64.
Writes gpurun_out/contraction_roofline.json (copied to profiles/)."""
import json, os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apyib_b200.contraction import contract
from apyib_b200._lib import lib, check

fl, ms = C.c_double(), C.c_float()
check(lib.apyib_peak_fp64(1, 8000, C.byref(fl), C.byref(ms)))
PEAK = fl.value / 1e12
out = {"fp64_dmma_peak_tflops": PEAK, "cases": []}


def rnd(shape, dt):
    x = torch.randn(shape, dtype=torch.float64, device="cuda")
    if dt == torch.complex128:
        x = torch.complex(x, torch.randn(shape, dtype=torch.float64, device="cuda"))
    return x


def run(name, spec, sa, sb, so, dt, reps=3):
    A, B, O = rnd(sa, dt), rnd(sb, dt), torch.zeros(so, dtype=dt, device="cuda")
    contract(spec, A, B, O, 1.0, 0.0)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    best = 1e30
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); contract(spec, A, B, O, 1.0, 1.0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    from apyib_b200.contraction import _plan
    M, N, K = _plan(spec, A, B, O)[:3]
    fac = 8.0 if dt == torch.complex128 else 2.0
    tf = fac * M * N * K / (best * 1e-3) / 1e12
    # spot check against torch on a slice (library used as checker only)
    rec = {"name": name, "spec": spec, "dtype": str(dt).split(".")[1], "M": M, "N": N, "K": K, "ms": best,
           "tflops": tf, "frac_of_dmma_peak": tf / PEAK}
    print(json.dumps(rec), flush=True)
    out["cases"].append(rec)
    del A, B, O
    torch.cuda.empty_cache()


c128, f64 = torch.complex128, torch.float64
quick = "--quick" in sys.argv
run("gemm 4096^3", "mk,kn->mn", (4096, 4096), (4096, 4096), (4096, 4096), f64)
run("gemm 4096^3", "mk,kn->mn", (4096, 4096), (4096, 4096), (4096, 4096), c128)
run("gemm A.B^T 2048^3", "mk,nk->mn", (2048, 2048), (2048, 2048), (2048, 2048), c128)
# methyloxirane/cc-pVDZ spatial (o=12 fc, v=70): ladder and ring
o, v = 12, 70
run("ladder spatial (12,70)", "abcd,ijcd->ijab", (v, v, v, v), (o, o, v, v), (o, o, v, v), c128)
run("ladder spatial (12,70)", "abcd,ijcd->ijab", (v, v, v, v), (o, o, v, v), (o, o, v, v), f64)
run("ring spatial (12,70)", "kbcj,ikac->ijab", (o, v, v, o), (o, o, v, v), (o, o, v, v), c128)
# config 5 sweep, spin-orbital: nso = 100 (O=20,V=80), 200 (O=40,V=160)
for O, V in ((20, 80),) + (() if quick else ((40, 160),)):
    run("ladder SO (%d,%d)" % (O, V), "abcd,ijcd->ijab", (V, V, V, V), (O, O, V, V), (O, O, V, V), c128)
    run("ring SO (%d,%d)" % (O, V), "kbcj,ikac->ijab", (O, V, V, O), (O, O, V, V), (O, O, V, V), c128)
    run("oooo SO (%d,%d)" % (O, V), "klij,klab->ijab", (O, O, O, O), (O, O, V, V), (O, O, V, V), c128)
# upper half of the config-5 sweep (--big): nso = 300 (O=60,V=240) whole, nso = 400 (O=80,V=320) as one eighth of the
# <ab||cd> rows (a < V/8): the full V^4 block is 168 GB complex128, and the ladder is row-separable, so the kernel
# sees the same N, K and tile stream; flops and time both cover that slice.
if "--big" in sys.argv:
    for O, V, frac in ((60, 240, 1), (80, 320, 8)):
        run("ladder SO (%d,%d)%s" % (O, V, "" if frac == 1 else " rows a<V/%d" % frac), "abcd,ijcd->ijab",
            (V // frac, V, V, V), (O, O, V, V), (O, O, V // frac, V), c128, reps=2)
        torch.cuda.empty_cache()
        run("ring SO (%d,%d)" % (O, V), "kbcj,ikac->ijab", (O, V, V, O), (O, O, V, V), (O, O, V, V), c128, reps=2)
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/contraction_roofline.json", "w"), indent=1)
