"""One AO->MO transform (utils.compute_ERI_MO_dev, four rotating-layout quarter transforms on the TMA-fed kernel) at nbf = 86,
4 frozen core orbitals -- target for `ncu --set full -k regex:contract_tma`."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apyib_b200
from apyib_b200.contraction import contract_new
nbf, nt = 86, 82
dt = torch.complex128 if os.environ.get("CPLX") else torch.float64
g = torch.Generator(device="cuda"); g.manual_seed(0)
G = torch.randn(nbf, nbf, nbf, nbf, dtype=torch.float64, device="cuda", generator=g).to(dt)
Ct = torch.randn(nt, nbf, dtype=torch.float64, device="cuda", generator=g).to(dt)
def run():
    X = contract_new("mnlg,sg->smnl", G, Ct)
    X = contract_new("smnl,rl->rsmn", X, Ct, conj_b=True)
    X = contract_new("rsmn,qn->qrsm", X, Ct)
    return contract_new("qrsm,pm->pqrs", X, Ct, conj_b=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = (8.0 if dt == torch.complex128 else 2.0) * sum(nbf ** (4 - k) * nt ** k * nt for k in range(4))
print("AO->MO transform nbf=%d nt=%d %s: %.3f ms, %.2f TFLOP/s" % (nbf, nt, dt, ms, fl / ms / 1e9))
