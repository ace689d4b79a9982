"""HBM roofline of the streaming kernels at (S)-methyloxirane/cc-pVDZ spin-orbital sizes
(O=24, V=140: n = 1.13e7 complex amplitudes = 181 MB per vector, well beyond the 126 MB L2).
Algorithmic bytes = arrays read + written.  Writes gpurun_out/streaming_roofline.json."""
import json, os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apyib_b200._lib import lib, check
from apyib_b200.device import ptr, stream_ptr, reduce_scratch, zeros, i32, i64
from apyib_b200.utils import gather4

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]
out = {"hbm_peak_gbs": PEAK, "cases": []}
O, V = 24, 140
n1, n2 = O * V, O * O * V * V
n = n1 + n2
c128 = torch.complex128
rnd = lambda *s: torch.complex(torch.randn(*s, dtype=torch.float64, device="cuda"), torch.randn(*s, dtype=torch.float64, device="cuda"))
scr = reduce_scratch()


def timeit(name, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    gbs = nbytes / (best * 1e-3) / 1e9
    rec = {"kernel": name, "ms": best, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / PEAK}
    print(json.dumps(rec), flush=True)
    out["cases"].append(rec)


r, t, told, w = rnd(n), rnd(n), rnd(n), rnd(n)
E = zeros((6,), torch.float64)
eps_o = torch.linspace(-2, -1, O // 2, dtype=torch.float64, device="cuda")
eps_v = torch.linspace(1, 3, V // 2, dtype=torch.float64, device="cuda")
timeit("copy (t_old = t)", lambda: check(lib.apyib_copy(1, ptr(told), ptr(t), n, stream_ptr())), 2 * n * 16)
timeit("ci_update (r -= E t; t += r/D)", lambda: check(lib.apyib_ci_update(1, ptr(r), ptr(t), ptr(E), ptr(eps_o), ptr(eps_v), O, V, 1, 1, 1, None, None, None, None, stream_ptr())), 4 * n * 16)
hist_e, hist_t = rnd(8, n), rnd(8, n)
B = zeros((128,), torch.float64); c = zeros((16,), torch.float64)
it = torch.full((1,), 8, dtype=torch.int32, device="cuda")
timeit("diis_push (m=8: copy r,t + 8 dots)", lambda: check(lib.apyib_diis_push(1, ptr(r), ptr(t), ptr(hist_e), ptr(hist_t), n, ptr(it), ptr(B), ptr(scr), 1, None, stream_ptr())), (2 + 2 + 7) * n * 16)
c[0] = 1.0
timeit("lincomb+energy+rms (m=8)", lambda: check(lib.apyib_lincomb_energy_rms(1, ptr(hist_t), n, 0, ptr(it), ptr(c), ptr(t), ptr(told), ptr(w), n1, n, ptr(E), ptr(scr), 1, None, stream_ptr())), (8 + 1 + 2) * n * 16)
o2 = zeros((2,), torch.float64)
timeit("dots (1 vector)", lambda: check(lib.apyib_dots(1, ptr(r), 0, 1, ptr(t), n, 1, ptr(o2), ptr(scr), stream_ptr())), 2 * n * 16)
timeit("axpby", lambda: check(lib.apyib_axpby(1, n, 0.5, 0.1, ptr(r), 1, 1.0, 0.0, ptr(t), stream_ptr())), 3 * n * 16)
# MP2 + gather4 on an MO tensor of methyloxirane spatial size (n = 82 active MOs, o = 12): 82^4 c128 = 723 MB
nmo, o = 82, 12
eri = rnd(nmo, nmo, nmo, nmo)
eps = torch.linspace(-2, 3, nmo, dtype=torch.float64, device="cuda")
v = nmo - o
t2 = torch.empty(o, o, v, v, dtype=c128, device="cuda")
timeit("mp2_t2_energy spatial (o=12,v=70)", lambda: check(lib.apyib_mp2_t2_energy(1, ptr(eri), nmo, o, ptr(eps), 2, ptr(t2), ptr(E), ptr(scr), None, stream_ptr())), 3 * o * o * v * v * 16)
t2so = torch.empty(2 * o, 2 * o, 2 * v, 2 * v, dtype=c128, device="cuda")
timeit("mp2_t2_energy spin-orbital (O=24,V=140)", lambda: check(lib.apyib_mp2_t2_energy(1, ptr(eri), nmo, o, ptr(eps), 3, ptr(t2so), ptr(E), ptr(scr), ptr(t2), stream_ptr())), (16 * o * o * v * v + 3 * o * o * v * v) * 16)
timeit("gather4 <ab||cd> SO block (V=140)", lambda: gather4(eri, 1, [2 * v] * 4, [0, 2, 1, 3], [2 * o] * 4, 1.0, [0, 3, 1, 2], [2 * o] * 4, -1.0), ((2 * v) ** 4 + 2 * v ** 4) * 16, reps=2)
del eri, t2, t2so
torch.cuda.empty_cache()
# config-5-like spatial MP2 (nso = 400: nbf = 200, o = 40, v = 160), real: source 12.8 GB, t2 0.33 GB
nmo, o = 200, 40
eri = torch.randn(nmo, nmo, nmo, nmo, dtype=torch.float64, device="cuda")
eps = torch.linspace(-2, 3, nmo, dtype=torch.float64, device="cuda")
v = nmo - o
t2 = torch.empty(o, o, v, v, dtype=torch.float64, device="cuda")
timeit("mp2_t2_energy spatial f64 (o=40,v=160)", lambda: check(lib.apyib_mp2_t2_energy(0, ptr(eri), nmo, o, ptr(eps), 2, ptr(t2), ptr(E), ptr(scr), None, stream_ptr())), 3 * o * o * v * v * 8)
t2so = torch.empty(2 * o, 2 * o, 2 * v, 2 * v, dtype=torch.float64, device="cuda")
timeit("mp2_t2_energy spin-orbital f64 (O=80,V=320)", lambda: check(lib.apyib_mp2_t2_energy(0, ptr(eri), nmo, o, ptr(eps), 3, ptr(t2so), ptr(E), ptr(scr), ptr(t2), stream_ptr())), (16 * o * o * v * v + 3 * o * o * v * v) * 8, reps=3)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/streaming_roofline.json", "w"), indent=1)
