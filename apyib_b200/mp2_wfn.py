"""MP2 wavefunction -- drop-in for apyib/mp2_wfn.py (same constructor, solve methods, returns).

The amplitudes and the energy come out of ONE streaming kernel (csrc/stream.cu: mp2_kernel):
denominators are formed on the fly from the orbital energies, the spin-orbital variant
spin-blocks and antisymmetrises the spatial MO integrals on the fly (the reference first
materialises the (2n)^4 spin-orbital tensor with an interpreted loop, utils.py:317-365).
"""
from __future__ import annotations

import numpy as np
import torch

from . import config
from ._lib import lib, check
from .device import to_device, to_host, empty, zeros, dtype_code, ptr, stream_ptr, reduce_scratch
from .utils import get_slices, compute_ERI_MO_dev


class mp2_wfn(object):
    """Reference: apyib/mp2_wfn.py:16-87."""

    def __init__(self, parameters, wfn):
        self.parameters = parameters
        self.H = wfn.H
        self.wfn = wfn
        self.C = wfn.C
        self.C_list, self.I_list = get_slices(self.parameters, self.wfn)
        o, v = self.C_list[1], self.C_list[2]
        self.eps_o = np.asarray(wfn.eps)[o]
        self.eps_v = np.asarray(wfn.eps)[v]
        self.D_ijab = (self.eps_o.reshape(-1, 1, 1, 1) + self.eps_o.reshape(-1, 1, 1)
                       - self.eps_v.reshape(-1, 1) - self.eps_v)           # mp2_wfn.py:39

    def _solve(self, spin_orbital):
        ERI_MO = compute_ERI_MO_dev(self.parameters, self.wfn, self.C_list)   # (n,n,n,n) chemists'
        n = ERI_MO.shape[0]
        o = len(self.eps_o)
        eps = to_device(np.concatenate([self.eps_o, self.eps_v]).real.astype(np.float64))
        O, V = (2 * o, 2 * (n - o)) if spin_orbital else (o, n - o)
        t2 = empty((O, O, V, V), ERI_MO.dtype)
        E = zeros((2,), torch.float64)
        work = empty((o, o, n - o, n - o), ERI_MO.dtype) if spin_orbital else None
        # bit 1: ERI_MO comes from compute_ERI_MO on real AO integrals, hence (pq|rs) = conj((qp|sr))
        check(lib.apyib_mp2_t2_energy(dtype_code(ERI_MO), ptr(ERI_MO), n, o, ptr(eps), int(spin_orbital) | 2,
                                      ptr(t2), ptr(E), ptr(reduce_scratch()), ptr(work), stream_ptr()))
        e = to_host(E)
        if ERI_MO.dtype == torch.complex128:
            E_MP2 = np.complex128(complex(e[0], e[1]))
        else:
            E_MP2 = np.float64(e[0])
        return E_MP2, (t2 if config.RETURN_DEVICE else to_host(t2))      # device-resident mode: bench `value` leg

    def solve_MP2(self):
        """Reference: mp2_wfn.py:42-59.  Returns (E_MP2, t2[o,o,v,v])."""
        return self._solve(False)

    def solve_MP2_SO(self):
        """Reference: mp2_wfn.py:64-87.  Returns (E_MP2, t2[O,O,V,V]) in the spin-orbital basis."""
        return self._solve(True)

    def perturbed_t2(self, t2, dF_MO, dERI, kind):
        """Closed-form perturbed MP2 amplitudes of the analytic AAT route:
            kind "H" (magnetic field,        analytic_aats.py:347-352)
            kind "R" (nuclear displacement,  analytic_aats.py:446-451)
        dt2 = [ d<ab|ij> (+/-) (dF . t2 terms) ] / D_ijab with `dERI` the perturbed PHYSICISTS' integrals d<pq|rs> and
        `dF_MO` the perturbed Fock matrix over the active MO space, exactly the reference's local variables
        `dERI_dH` / `dERI_dR`, `df_dH` / `df_dR` (host inputs from CPHF + Psi4 derivative integrals).  The four
        contractions run on the DMMA contraction kernel, the division on the Jacobi-update kernel.  Returns dt2."""
        from .ci_wfn import _Engine
        from .contraction import contract
        from .utils import gather4
        O, V = len(self.eps_o), len(self.eps_v)
        cplx = any(np.iscomplexobj(x) for x in (t2, dF_MO, dERI))
        dt = torch.complex128 if cplx else torch.float64
        T2, dF, dW = (to_device(np.asarray(x), dt) if not isinstance(x, torch.Tensor) else to_device(x, dt) for x in (t2, dF_MO, dERI))
        eng = _Engine({"DIIS": False}, dt, 1, O, V, False, False, self.eps_o[None, :], self.eps_v[None, :])
        r = eng.t2(eng.r)[0]
        Foo, Fvv = dF[:O, :O], dF[O:, O:]
        if kind == "H":
            gather4(dW, 0, (O, O, V, V), [2, 3, 0, 1], [O, O, 0, 0], out=r)        # dERI.swapaxes(0,2).swapaxes(1,3)[o,o,v,v]
            contract("ac,ijcb->ijab", Fvv, T2, r, 1.0, 1.0)
            contract("bc,ijac->ijab", Fvv, T2, r, 1.0, 1.0)
            contract("ki,kjab->ijab", Foo, T2, r, -1.0, 1.0)
            contract("kj,ikab->ijab", Foo, T2, r, -1.0, 1.0)
        elif kind == "R":
            gather4(dW, 0, (O, O, V, V), [0, 1, 2, 3], [0, 0, O, O], out=r)        # dERI[o,o,v,v]
            contract("kjab,ik->ijab", T2, Foo, r, -1.0, 1.0)
            contract("ikab,kj->ijab", T2, Foo, r, -1.0, 1.0)
            contract("ijcb,ac->ijab", T2, Fvv, r, 1.0, 1.0)
            contract("ijac,cb->ijab", T2, Fvv, r, 1.0, 1.0)
        else:
            raise ValueError("kind must be 'H' or 'R'")
        eng.out6.zero_()
        eng.t.zero_()
        eng._update()                                    # E = 0, t = 0:  t <- r / D_ijab
        out = eng.t2()[0]
        return out.clone() if config.RETURN_DEVICE else to_host(out)
