import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apyib_b200.contraction import contract
n = int(os.environ.get("N", 2048))
for dt in (torch.float64, torch.complex128):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda").to(dt)
    B = torch.randn(n, n, dtype=torch.float64, device="cuda").to(dt)
    O = torch.zeros(n, n, dtype=dt, device="cuda")
    for _ in range(2):
        contract("mk,kn->mn", A, B, O, 1.0, 0.0)
    torch.cuda.synchronize()
