"""Times the six ring terms + ladder + thin terms of the spatial CISD residual (ci_wfn.py:457-482) at the
(S)-methyloxirane/cc-pVDZ frozen-core shape, batched over NB points -- target for ncu -k regex:contract."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apyib_b200.contraction import contract
o, v, nb = int(os.environ.get("O", 12)), int(os.environ.get("V", 70)), int(os.environ.get("NB", 61))
dt = torch.complex128 if os.environ.get("CPLX") else torch.float64
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=g).to(dt)
t1, t2 = rnd(nb, o, v), rnd(nb, o, o, v, v)
r1, r2 = torch.zeros_like(t1), torch.zeros_like(t2)
W = dict(ovvo=rnd(nb, o, v, o, v), ovov=rnd(nb, o, v, o, v), vvvo=rnd(nb, o, v, v, v), vvov=rnd(nb, o, v, v, v),
         ovoo=rnd(nb, o, o, o, v), vooo=rnd(nb, o, o, o, v), oooo=rnd(nb, o, o, o, o), vovv=rnd(nb, v, o, v, v),
         ooov=rnd(nb, o, o, o, v), Fvv=rnd(nb, v, v), Foo=rnd(nb, o, o), Fov=rnd(nb, o, v))
if not os.environ.get("NOLADDER"):
    W["vvvv"] = rnd(nb, v, v, v, v)
terms = [("skbjc,sikca->sijab", "ovvo", "t2"), ("skaic,skjcb->sijab", "ovvo", "t2"), ("skbic,skjac->sijab", "ovov", "t2"),
         ("skaic,skjbc->sijab", "ovvo", "t2"), ("skbjc,sikac->sijab", "ovvo", "t2"), ("skajc,sikcb->sijab", "ovov", "t2"),
         ("sabcd,sijcd->sijab", "vvvv", "t2"), ("sklij,sklab->sijab", "oooo", "t2"),
         ("sjabc,sic->sijab", "vvvo", "t1"), ("siabc,sjc->sijab", "vvov", "t1"), ("skijb,ska->sijab", "ovoo", "t1"),
         ("skija,skb->sijab", "vooo", "t1"), ("sac,sijcb->sijab", "Fvv", "t2"), ("sbc,sijac->sijab", "Fvv", "t2"),
         ("ski,skjab->sijab", "Foo", "t2"), ("skj,sikab->sijab", "Foo", "t2"),
         ("sajbc,sijbc->sia", "vovv", "t2"), ("skjib,skjab->sia", "ooov", "t2"), ("sjb,sijab->sia", "Fov", "t2")]
if "vvvv" in W:        # pair-packed ladder (ci_wfn._PackedLadder): N = o(o+1)/2 occupied pairs
    npair = o * (o + 1) // 2
    W["tp"], W["hp"] = rnd(nb, npair, v, v), torch.zeros(nb, npair, v, v, dtype=dt, device=dev)
    terms.append(("sabcd,spcd->spab", "vvvv", "tp"))
only = os.environ.get("ONLY")
flop_unit = 8.0 if dt == torch.complex128 else 2.0
for spec, wk, tk in terms:
    if wk not in W or (only and only not in spec):
        continue
    A, B = W[wk], (t2 if tk == "t2" else (W["tp"] if tk == "tp" else t1))
    out = r1 if spec.endswith("sia") else (W["hp"] if tk == "tp" else r2)
    for _ in range(2):
        contract(spec, A, B, out, 1.0, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        contract(spec, A, B, out, 1.0, 1.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ins = spec.split("->")[0].split(",")
    idx = set(ins[0]) | set(ins[1])
    size = dict(s=nb, i=o, j=o, k=o, l=o, a=v, b=v, c=v, d=v, p=o * (o + 1) // 2)
    fl = flop_unit * np.prod([float(size[c]) for c in idx])
    byts = (A.numel() + B.numel() + 2 * out.numel()) * A.element_size()
    print("%-22s %8.3f ms  %6.2f TFLOP/s  %7.1f GB/s (operands once)" % (spec, ms, fl / ms / 1e9, byts / ms / 1e6))
