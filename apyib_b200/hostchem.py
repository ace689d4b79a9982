"""Host-side inputs of the hot path (north star: AO integrals and the complex-HF SCF stay on the host).

The numpy implementation lives in the neutral top-level package `hostinputs` (the oracle and bench.py's reference
arm use it too, without importing the product); this module re-exports it under its old name and adds the one
piece that touches the device: the optional (2J - K)[D] accelerator of the host SCF.
"""
from __future__ import annotations

import os as _os
import sys as _sys

import numpy as np

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:                                  # `hostinputs` sits next to this package in the repo
    _sys.path.insert(0, _root)

from hostinputs.chem import *                               # noqa: F401,F403
from hostinputs.chem import (Molecule, BasisHandle, provider_ao_overlap, SGaussianProvider, SyntheticProvider,   # noqa: F401
                             Psi4Provider, select_provider, Hamiltonian, hf_wfn, JK_HOOK, BOHR2ANG, _diis)


def _device_jk(wfn):
    """config.SCF_DEVICE_JK: the nbf^4 part of the Fock build, (2J - K)[D] = sum_ls D_ls (2 (mn|ls) - (ml|ns))
    (hf_wfn.py:81-107, utils.py:249), as one launch of the contraction kernel per SCF iteration on a device copy of
    2(mn|ls) - (ml|ns) (one gather from the AO integrals, which the correlated solver needs on the device anyway).
    D goes up and (2J - K)[D] comes back as nbf^2 numbers; DIIS, the nbf^3 algebra and the eigensolver stay on the
    host.  Real AO integrals only (returns None otherwise -> numpy path)."""
    from . import config
    if not config.SCF_DEVICE_JK:
        return None
    import torch
    if not torch.cuda.is_available():
        return None
    from .contraction import contract_new
    from .device import to_device, to_host
    from .utils import gather4
    H = wfn.H
    if np.iscomplexobj(H.ERI):
        return None
    n = wfn.nbf
    cache = H.__dict__.setdefault("_apyib_b200_dev", None) or {}
    H._apyib_b200_dev = cache
    if "r" in cache:
        G = cache["r"][1]
    else:
        G = cache.get("eri_r")
        if G is None:
            G = cache["eri_r"] = to_device(np.asarray(H.ERI), torch.float64)
    GK = gather4(G, 0, (n, n, n, n), [0, 1, 2, 3], [0, 0, 0, 0], 2.0, [0, 2, 1, 3], [0, 0, 0, 0], -1.0).reshape(n * n, n * n)

    def jk(d):
        parts = np.stack([d.real, d.imag]) if np.iscomplexobj(d) else d.reshape(1, -1)
        out = to_host(contract_new("xy,qy->qx", GK, to_device(np.ascontiguousarray(parts), torch.float64)))
        return out[0] + 1j * out[1] if np.iscomplexobj(d) else out[0]

    return jk


JK_HOOK[0] = _device_jk
