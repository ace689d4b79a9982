"""FP64 / HBM calibration on the GPU box -> gpurun_out/calibration.json (copied to profiles/)."""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apyib_b200._lib import lib, check

out = {"gpu": torch.cuda.get_device_name(0)}
for name, flag in (("dmma", 1), ("dfma", 0)):
    fl, ms = C.c_double(), C.c_float()
    check(lib.apyib_peak_fp64(flag, 20000, C.byref(fl), C.byref(ms)))
    out["fp64_%s_tflops" % name] = fl.value / 1e12
    out["fp64_%s_ms" % name] = ms.value
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")   # 2 GiB
b = torch.empty_like(a)
bw = C.c_double()
check(lib.apyib_peak_copy(C.c_void_p(b.data_ptr()), C.c_void_p(a.data_ptr()), a.numel() * 8, 5, C.byref(bw)))
out["copy_gbs_own_kernel"] = bw.value / 1e9
del a, b
# cuBLAS references (library calibration only, never on the product path)
for dt, nm, fac in ((torch.float64, "cublas_dgemm_tflops", 2.0), (torch.complex128, "cublas_zgemm_real_tflops", 8.0)):
    n = 4096
    x = torch.randn(n, n, dtype=dt, device="cuda"); y = torch.randn(n, n, dtype=dt, device="cuda")
    for _ in range(2): x @ y
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); x @ y; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[nm] = fac * n ** 3 / (best * 1e-3) / 1e12
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/calibration.json", "w"), indent=1)
print(json.dumps(out))
