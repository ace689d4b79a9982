#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "target_shape" --durations=10 > gpurun_out/t_target.log 2>&1
echo "target tests rc=$?"; tail -25 gpurun_out/t_target.log
python - <<'PY'
import torch, time, numpy as np
n = 437*1024*1024//8
a = torch.empty(n, dtype=torch.float64).pin_memory()
b = torch.empty(n, dtype=torch.float64)          # pageable
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, src in (("pinned", a), ("pageable", b)):
    for _ in range(2):
        torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, "H2D GB/s", n*8/dt/1e9)
torch.cuda.synchronize(); t=time.perf_counter(); a.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("D2H pinned GB/s", n*8/(time.perf_counter()-t)/1e9)
t=time.perf_counter(); a.copy_(b); print("host memcpy pageable->pinned GB/s (1 thread)", n*8/(time.perf_counter()-t)/1e9)
# cudaHostRegister in place on a numpy array
x = np.empty(n); 
t=time.perf_counter(); rc = torch.cuda.cudart().cudaHostRegister(x.ctypes.data, x.nbytes, 0); print("register rc", rc, time.perf_counter()-t)
tx = torch.from_numpy(x); print("is_pinned after register:", tx.is_pinned())
torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(tx, non_blocking=True); torch.cuda.synchronize(); print("registered numpy H2D GB/s", n*8/(time.perf_counter()-t)/1e9)
import os; print("cpus", os.cpu_count())
PY
