"""CID / CISD wavefunctions -- drop-in for apyib/ci_wfn.py.

Same constructor (AO->MO transform and MO Fock build happen here, ci_wfn.py:46-47), same four
solve methods and return shapes.  Everything between "integrals are on the device" and
"converged amplitudes come back" runs in libapyib_b200 kernels:

  * every `oe.contract(...)` line of the reference's residuals is one launch of the DMMA
    contraction kernel (index regrouping via offset tables, no transposed copies);
  * `r -= E t; t += r/D`, the DIIS Gram row, the bordered solve, the extrapolation, the new energy
    and both rms sums are streaming / tiny kernels with deterministic reductions;
  * the iteration counter, energy and DIIS state live on the device, so one captured CUDA graph
    replays every iteration; the host only reads back 6 doubles per iteration for the
    convergence test (which keeps the reference's exact semantics, ci_wfn.py:123-130).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import config
from ._lib import lib, check
from .contraction import contract
from .device import (to_device, to_host, empty, zeros, dtype_code, ptr, stream_ptr, reduce_scratch, i32, i64,
                     Graph)
from .utils import (get_slices, compute_F_MO_dev, compute_ERI_MO_dev, spin_block_2_dev, gather4)

_NULL = C.c_void_p(0)


def w_block(E, labels, out_labels, space, bounds, spin=0, c1=1.0, c2=0.0):
    """Dense block of the physicists' integrals W[p,q,r,s] = (pr|qs) built from the chemists'
    tensor E (optionally spin-blocked on the fly):
        out[out_labels] = c1 * W[labels] + c2 * W[labels with the last two swapped]
    `space[label]` in 'ov' picks the index range from `bounds` = {'o': (0, O), 'v': (O, O+V)}."""
    l0, l1, l2, l3 = labels
    src1 = (l0, l2, l1, l3)            # W[l0,l1,l2,l3] = E[l0,l2,l1,l3]
    src2 = (l0, l3, l1, l2)            # W[l0,l1,l3,l2] = E[l0,l3,l1,l2]
    shape = [bounds[space[ch]][1] - bounds[space[ch]][0] for ch in out_labels]
    perm = lambda src: [out_labels.index(ch) for ch in src]
    start = lambda src: [bounds[space[ch]][0] for ch in src]
    return gather4(E, spin, shape, perm(src1), start(src1), c1, perm(src2), start(src2), c2)


class _Engine:
    """Device-resident Jacobi/DIIS iteration shared by the four solvers."""

    def __init__(self, parameters, dtype, O, V, has_singles, spin_orbital, eps_o, eps_v, symmetrize=False):
        self.p = parameters
        self.dtype, self.O, self.V = dtype, O, V
        self.n1 = O * V if has_singles else 0
        self.n2 = O * O * V * V
        self.len = self.n1 + self.n2
        self.has_singles, self.so, self.symmetrize = has_singles, int(spin_orbital), symmetrize
        self.code = 1 if dtype == torch.complex128 else 0
        self.eps_o = to_device(np.asarray(eps_o).real.astype(np.float64))
        self.eps_v = to_device(np.asarray(eps_v).real.astype(np.float64))
        z = lambda n: zeros((n,), dtype)
        self.t, self.r, self.t_old, self.w, self.r0 = z(self.len), z(self.len), z(self.len), z(self.len), z(self.len)
        self.r_half = z(self.n2) if symmetrize else None
        self.out6 = zeros((6,), torch.float64)        # E(re,im), S1(re,im), S2(re,im)
        self.diis = bool(parameters["DIIS"])
        if self.diis:
            self.hist_e = zeros((8, self.len), dtype)
            self.hist_t = zeros((8, self.len), dtype)
            self.B = zeros((8 * 8 * 2,), torch.float64)
            self.c = zeros((16,), torch.float64)
            self.iter = torch.ones((1,), dtype=torch.int32, device=self.t.device)
        self.scratch = reduce_scratch()

    # views into the concatenated vectors
    def t1(self, x=None):
        x = self.t if x is None else x
        return x[:self.n1].view(self.O, self.V)

    def t2(self, x=None):
        x = self.t if x is None else x
        return x[self.n1:].view(self.O, self.O, self.V, self.V)

    def _update(self):
        check(lib.apyib_ci_update(self.code, ptr(self.r), ptr(self.t), ptr(self.out6), ptr(self.eps_o),
                                  ptr(self.eps_v), self.O, self.V, int(self.has_singles), self.so, stream_ptr()))

    def _energy_rms(self, use_diis):
        if use_diis:
            check(lib.apyib_lincomb_energy_rms(self.code, ptr(self.hist_t), self.len, 0, ptr(self.iter), ptr(self.c),
                                               ptr(self.t), ptr(self.t_old), ptr(self.w), self.n1, self.len,
                                               ptr(self.out6), ptr(self.scratch), stream_ptr()))
        else:
            check(lib.apyib_lincomb_energy_rms(self.code, _NULL, 0, 0, _NULL, _NULL, ptr(self.t), ptr(self.t_old),
                                               ptr(self.w), self.n1, self.len, ptr(self.out6), ptr(self.scratch),
                                               stream_ptr()))

    def _copy(self, dst, src, n):
        check(lib.apyib_copy(self.code, ptr(dst), ptr(src), n, stream_ptr()))

    def initial_guess(self):
        """t = r0 / D, E = w.t   (ci_wfn.py:66-70, 193-197, 286-293, 435-442)."""
        self.out6.zero_()
        self.t.zero_()
        self._copy(self.r, self.r0, self.len)
        if self.symmetrize:                   # CID spatial keeps 0.5*K in r0 (ci_wfn.py:83); guess uses K
            check(lib.apyib_axpby(self.code, self.n2, 2.0, 0.0, ptr(self.r0[self.n1:]), 0, 0.0, 0.0,
                                  ptr(self.r[self.n1:]), stream_ptr()))
        self._update()
        self._copy(self.t_old, self.t, self.len)
        self._energy_rms(False)

    def iteration(self, residual):
        """One trip of the reference's while-loop body, enqueued on the current stream."""
        self._copy(self.t_old, self.t, self.len)
        if self.symmetrize:
            self._copy(self.r[:self.n1], self.r0[:self.n1], self.n1)
            self._copy(self.r_half, self.r0[self.n1:], self.n2)
            residual(self.r_half)
            check(lib.apyib_symmetrize_ijab(self.code, ptr(self.r_half), ptr(self.r[self.n1:]), self.O, self.V,
                                            stream_ptr()))
        else:
            self._copy(self.r, self.r0, self.len)
            residual(None)
        self._update()
        if self.diis:
            check(lib.apyib_diis_push(self.code, ptr(self.r), ptr(self.t), ptr(self.hist_e), ptr(self.hist_t),
                                      self.len, ptr(self.iter), ptr(self.B), ptr(self.scratch), stream_ptr()))
            check(lib.apyib_diis_solve(self.code, ptr(self.B), 8, 0, ptr(self.iter), ptr(self.c), stream_ptr()))
            self._energy_rms(True)
            check(lib.apyib_iter_advance(ptr(self.iter), stream_ptr()))
        else:
            self._energy_rms(False)

    def read(self):
        h = to_host(self.out6)
        if self.code:
            return (np.complex128(complex(h[0], h[1])), np.complex128(complex(h[2], h[3])),
                    np.complex128(complex(h[4], h[5])))
        return np.float64(h[0]), np.float64(h[2]), np.float64(h[4])

    def run(self, residual, print_level=0, E_SCF=0.0, E_nuc=0.0):
        """The reference's iteration control (ci_wfn.py:76-130 etc.), verbatim semantics."""
        p = self.p
        self.initial_guess()
        E, _, _ = self.read()
        graph = None
        iteration = 1
        self.iterations = 0
        while iteration <= p["max_iterations"]:
            E_old = E
            if iteration == 1 or not config.USE_CUDA_GRAPH:
                self.iteration(residual)              # eager: also builds/caches the offset tables
            else:
                if graph is None:
                    graph = Graph()
                    graph.capture(lambda: self.iteration(residual))
                graph.launch()
            E, S1, S2 = self.read()
            rms_t1, rms_t2 = np.sqrt(S1), np.sqrt(S2)
            delta_E = E_old - E
            self.iterations = iteration
            if print_level > 0:
                E_tot = E_SCF + E + E_nuc
                print(" %02d %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f" % (
                    iteration, np.real(E), np.imag(E), np.real(E_tot), np.real(delta_E), np.imag(delta_E),
                    np.real(rms_t1), np.imag(rms_t1), np.real(rms_t2), np.imag(rms_t2)))
            if iteration > 1:
                conv = abs(delta_E) < p["e_convergence"] and rms_t2 < p["d_convergence"]
                if self.has_singles:
                    conv = conv and rms_t1 < p["d_convergence"]
                if conv:
                    break
            if iteration == p["max_iterations"]:
                if abs(delta_E) > p["e_convergence"] or rms_t2 > p["d_convergence"]:
                    print("Not converged.")
            iteration += 1
        return E


class ci_wfn(object):
    """Reference: apyib/ci_wfn.py:19-575."""

    def __init__(self, parameters, wfn):
        self.parameters = parameters
        self.H = wfn.H
        self.wfn = wfn
        self.C = wfn.C
        self.C_list, self.I_list = get_slices(self.parameters, self.wfn)
        o, v = self.C_list[1], self.C_list[2]
        self.eps_o = np.asarray(wfn.eps)[o]
        self.eps_v = np.asarray(wfn.eps)[v]
        self.D_ia = self.eps_o.reshape(-1, 1) - self.eps_v                                     # ci_wfn.py:42
        self.D_ijab = (self.eps_o.reshape(-1, 1, 1, 1) + self.eps_o.reshape(-1, 1, 1)
                       - self.eps_v.reshape(-1, 1) - self.eps_v)                               # ci_wfn.py:43
        self._F_dev, self.E_fc = compute_F_MO_dev(self.parameters, self.wfn, self.C_list)      # ci_wfn.py:46
        self._ERI_dev = compute_ERI_MO_dev(self.parameters, self.wfn, self.C_list)             # ci_wfn.py:47
        self._F_host = self._ERI_host = None
        self.iterations = 0

    # numpy views of the MO integrals, as the reference exposes them (analytic_aats.py reads these)
    @property
    def F_MO(self):
        if self._F_host is None:
            self._F_host = to_host(self._F_dev)
        return self._F_host

    @property
    def ERI_MO(self):
        if self._ERI_host is None:
            self._ERI_host = to_host(self._ERI_dev)
        return self._ERI_host

    # ---------------------------------------------------------------------------------------
    def _sizes(self, so):
        n = self._F_dev.shape[0]
        o = len(self.eps_o)
        f = 2 if so else 1
        O, V = f * o, f * (n - o)
        return O, V, {"o": (0, O), "v": (O, O + V)}

    def _finish(self, eng, E, singles):
        self.iterations = eng.iterations
        if config.RETURN_DEVICE:
            t2 = eng.t2().clone()
            return (E, eng.t1().clone(), t2) if singles else (E, t2)
        t2 = to_host(eng.t2()).copy()
        if singles:
            return E, to_host(eng.t1()).copy(), t2
        return E, t2

    # ---------------------------------------------------------------------------------------
    def solve_CID(self, print_level=0):
        """Spatial-orbital CID (ci_wfn.py:51-167).  Returns (E_CID, t2)."""
        O, V, bd = self._sizes(False)
        E4, F = self._ERI_dev, self._F_dev
        sp = dict(i="o", j="o", m="o", n="o", a="v", b="v", e="v", f="v")
        eng = _Engine(self.parameters, E4.dtype, O, V, False, False, self.eps_o, self.eps_v, symmetrize=True)
        blk = lambda lab, out=None, c1=1.0, c2=0.0: w_block(E4, lab, out or lab, sp, bd, 0, c1, c2)
        eng.r0.copy_(blk("abij", "ijab", 0.5).reshape(-1))                  # 0.5 <ab|ij>     ci_wfn.py:83
        eng.w.copy_(blk("ijab", None, 2.0, -1.0).reshape(-1))               # 2<ij|ab>-<ij|ba> ci_wfn.py:70
        Woooo, Wvvvv = blk("mnij"), blk("abef")
        Wovvo, Wovov = blk("mbej"), blk("mbie")
        Lovvo = blk("mbej", None, 1.0, -1.0)                                # <mb|ej>-<mb|je>  ci_wfn.py:89
        Foo, Fvv = F[:O, :O], F[O:, O:]

        def residual(r):
            t2 = eng.t2()
            r = r.view(O, O, V, V)
            contract("ijae,be->ijab", t2, Fvv, r, 1.0, 1.0)                 # ci_wfn.py:84
            contract("imab,mj->ijab", t2, Foo, r, -1.0, 1.0)                # :85
            contract("mnab,mnij->ijab", t2, Woooo, r, 0.5, 1.0)             # :86
            contract("ijef,abef->ijab", t2, Wvvvv, r, 0.5, 1.0)             # :87
            contract("imae,mbej->ijab", t2, Wovvo, r, 1.0, 1.0)             # :88  (t2 - t2.swapaxes(2,3)) . W
            contract("imea,mbej->ijab", t2, Wovvo, r, -1.0, 1.0)
            contract("imae,mbej->ijab", t2, Lovvo, r, 1.0, 1.0)             # :89
            contract("mjae,mbie->ijab", t2, Wovov, r, -1.0, 1.0)            # :90

        E = eng.run(residual, print_level, self.wfn.E_SCF, self.H.E_nuc)
        out = self._finish(eng, E, False)
        if config.VERBOSE and not config.RETURN_DEVICE:
            print("t-Amplitude Data:")
            print("Maximum t2: ", np.max(out[1]))
        return out

    # ---------------------------------------------------------------------------------------
    def solve_CID_SO(self, print_level=0):
        """Spin-orbital CID (ci_wfn.py:171-259).  Returns (E_CID, t2) in the spin-orbital basis."""
        O, V, bd = self._sizes(True)
        E4 = self._ERI_dev
        F = spin_block_2_dev(self._F_dev)                                   # compute_F_SO, ci_wfn.py:185
        sp = dict(i="o", j="o", m="o", n="o", a="v", b="v", e="v", f="v")
        eng = _Engine(self.parameters, E4.dtype, O, V, False, True, self.eps_o, self.eps_v)
        A = lambda lab, out=None, s=1.0: w_block(E4, lab, out or lab, sp, bd, 1, s, -s)     # <pq||rs>
        eng.r0.copy_(A("abij", "ijab").reshape(-1))                         # ci_wfn.py:210
        eng.w.copy_(A("ijab", None, 0.25).reshape(-1))                      # ci_wfn.py:197
        Aoooo, Avvvv, Aovvo = A("mnij"), A("abef"), A("mbej")
        Foo, Fvv = F[:O, :O], F[O:, O:]

        def residual(_):
            t2, r = eng.t2(), eng.t2(eng.r)
            contract("ijae,be->ijab", t2, Fvv, r, 1.0, 1.0)                 # :211
            contract("ijeb,ae->ijab", t2, Fvv, r, 1.0, 1.0)
            contract("imab,mj->ijab", t2, Foo, r, -1.0, 1.0)                # :212
            contract("mjab,mi->ijab", t2, Foo, r, -1.0, 1.0)
            contract("mnab,mnij->ijab", t2, Aoooo, r, 0.5, 1.0)             # :213
            contract("ijef,abef->ijab", t2, Avvvv, r, 0.5, 1.0)             # :214
            contract("imae,mbej->ijab", t2, Aovvo, r, 1.0, 1.0)             # :215
            contract("mjae,mbei->ijab", t2, Aovvo, r, 1.0, 1.0)             # :216
            contract("imeb,maej->ijab", t2, Aovvo, r, 1.0, 1.0)             # :217
            contract("mjeb,maei->ijab", t2, Aovvo, r, 1.0, 1.0)             # :218

        E = eng.run(residual, print_level, self.wfn.E_SCF, self.H.E_nuc)
        return self._finish(eng, E, False)

    # ---------------------------------------------------------------------------------------
    def solve_CISD_SO(self, print_level=0):
        """Spin-orbital CISD (ci_wfn.py:263-416).  Returns (E_CISD, t1, t2)."""
        O, V, bd = self._sizes(True)
        E4 = self._ERI_dev
        F = spin_block_2_dev(self._F_dev)
        sp = dict(i="o", j="o", k="o", l="o", a="v", b="v", c="v", d="v")
        eng = _Engine(self.parameters, E4.dtype, O, V, True, True, self.eps_o, self.eps_v)
        A = lambda lab, out=None, s=1.0: w_block(E4, lab, out or lab, sp, bd, 1, s, -s)
        n1 = eng.n1
        Foo, Fvv, Fov = F[:O, :O], F[O:, O:], F[:O, O:]
        eng.r0[:n1].copy_(F[O:, :O].transpose(0, 1).reshape(-1))            # F_ai as [i,a]   ci_wfn.py:309
        eng.r0[n1:].copy_(A("abij", "ijab").reshape(-1))                    # :319
        eng.w[:n1].copy_(Fov.reshape(-1))                                   # :355
        eng.w[n1:].copy_(A("ijab", None, 0.25).reshape(-1))
        Aovvo, Avovv, Aooov = A("jabi"), A("ajcb"), A("kjib")
        Aovoo, Avooo, Avvvo, Avvov = A("kbij"), A("akij"), A("abcj"), A("abic")
        Aoooo, Avvvv = A("klij"), A("abcd")
        Fov_c = Fov.contiguous()

        def residual(_):
            t1, t2 = eng.t1(), eng.t2()
            r1, r2 = eng.t1(eng.r), eng.t2(eng.r)
            contract("ji,ja->ia", Foo, t1, r1, -1.0, 1.0)                   # :310
            contract("ab,ib->ia", Fvv, t1, r1, 1.0, 1.0)                    # :311
            contract("jabi,jb->ia", Aovvo, t1, r1, 1.0, 1.0)                # :312
            contract("jb,ijab->ia", Fov_c, t2, r1, 1.0, 1.0)                # :313
            contract("ajcb,ijcb->ia", Avovv, t2, r1, 0.5, 1.0)              # :314
            contract("kjib,kjab->ia", Aooov, t2, r1, -0.5, 1.0)             # :315
            contract("kbij,ka->ijab", Aovoo, t1, r2, -1.0, 1.0)             # :320
            contract("akij,kb->ijab", Avooo, t1, r2, -1.0, 1.0)             # :321
            contract("abcj,ic->ijab", Avvvo, t1, r2, 1.0, 1.0)              # :322
            contract("abic,jc->ijab", Avvov, t1, r2, 1.0, 1.0)              # :323
            contract("bc,ijac->ijab", Fvv, t2, r2, 1.0, 1.0)                # :324
            contract("ac,ijcb->ijab", Fvv, t2, r2, 1.0, 1.0)                # :325
            contract("kj,ikab->ijab", Foo, t2, r2, -1.0, 1.0)               # :326
            contract("ki,kjab->ijab", Foo, t2, r2, -1.0, 1.0)               # :327
            contract("klij,klab->ijab", Aoooo, t2, r2, 0.5, 1.0)            # :328
            contract("abcd,ijcd->ijab", Avvvv, t2, r2, 0.5, 1.0)            # :329
            contract("kbcj,ikac->ijab", Aovvo, t2, r2, 1.0, 1.0)            # :330
            contract("kbci,kjac->ijab", Aovvo, t2, r2, 1.0, 1.0)            # :331
            contract("kacj,ikcb->ijab", Aovvo, t2, r2, 1.0, 1.0)            # :332
            contract("kaci,kjcb->ijab", Aovvo, t2, r2, 1.0, 1.0)            # :333

        E = eng.run(residual, print_level, self.wfn.E_SCF, self.H.E_nuc)
        return self._finish(eng, E, True)

    # ---------------------------------------------------------------------------------------
    def solve_CISD(self, print_level=0):
        """Spatial-orbital CISD (ci_wfn.py:420-574).  Returns (E_CISD, t1, t2)."""
        O, V, bd = self._sizes(False)
        E4, F = self._ERI_dev, self._F_dev
        sp = dict(i="o", j="o", k="o", l="o", a="v", b="v", c="v", d="v")
        eng = _Engine(self.parameters, E4.dtype, O, V, True, False, self.eps_o, self.eps_v)
        blk = lambda lab, out=None, c1=1.0, c2=0.0: w_block(E4, lab, out or lab, sp, bd, 0, c1, c2)
        n1 = eng.n1
        Foo, Fvv, Fov = F[:O, :O], F[O:, O:], F[:O, O:].contiguous()
        eng.r0[:n1].copy_(F[O:, :O].transpose(0, 1).reshape(-1))            # ci_wfn.py:457
        eng.r0[n1:].copy_(blk("abij", "ijab").reshape(-1))                  # :466
        check(lib.apyib_axpby(eng.code, n1, 2.0, 0.0, ptr(Fov), 0, 0.0, 0.0, ptr(eng.w), stream_ptr()))  # 2 F_ov  :504
        eng.w[n1:].copy_(blk("ijab", None, 2.0, -1.0).reshape(-1))
        Wovvo, Wovov = blk("kbcj"), blk("kbic")
        Lovvo = blk("jabi", None, 2.0, -1.0)                                # 2<ja|bi>-<ja|ib>  :460,478,481
        Lvovv = blk("ajbc", None, 2.0, -1.0)                                # :462
        Looov = blk("kjib", None, 2.0, -1.0)                                # :463
        Wvvvo, Wvvov, Wovoo, Wvooo = blk("abcj"), blk("abic"), blk("kbij"), blk("akij")
        Woooo, Wvvvv = blk("klij"), blk("abcd")

        def residual(_):
            t1, t2 = eng.t1(), eng.t2()
            r1, r2 = eng.t1(eng.r), eng.t2(eng.r)
            contract("ji,ja->ia", Foo, t1, r1, -1.0, 1.0)                   # :458
            contract("ab,ib->ia", Fvv, t1, r1, 1.0, 1.0)                    # :459
            contract("jabi,jb->ia", Lovvo, t1, r1, 1.0, 1.0)                # :460
            contract("jb,ijab->ia", Fov, t2, r1, 2.0, 1.0)                  # :461  F.(2 t2 - t2^T)
            contract("jb,ijba->ia", Fov, t2, r1, -1.0, 1.0)
            contract("ajbc,ijbc->ia", Lvovv, t2, r1, 1.0, 1.0)              # :462
            contract("kjib,kjab->ia", Looov, t2, r1, -1.0, 1.0)             # :463
            contract("abcj,ic->ijab", Wvvvo, t1, r2, 1.0, 1.0)              # :467
            contract("abic,jc->ijab", Wvvov, t1, r2, 1.0, 1.0)              # :468
            contract("kbij,ka->ijab", Wovoo, t1, r2, -1.0, 1.0)             # :469
            contract("akij,kb->ijab", Wvooo, t1, r2, -1.0, 1.0)             # :470
            contract("ac,ijcb->ijab", Fvv, t2, r2, 1.0, 1.0)                # :471
            contract("bc,ijac->ijab", Fvv, t2, r2, 1.0, 1.0)                # :472
            contract("ki,kjab->ijab", Foo, t2, r2, -1.0, 1.0)               # :473
            contract("kj,ikab->ijab", Foo, t2, r2, -1.0, 1.0)               # :474
            contract("klij,klab->ijab", Woooo, t2, r2, 1.0, 1.0)            # :475
            contract("abcd,ijcd->ijab", Wvvvv, t2, r2, 1.0, 1.0)            # :476
            contract("kbcj,ikca->ijab", Wovvo, t2, r2, -1.0, 1.0)           # :477
            contract("kaci,kjcb->ijab", Lovvo, t2, r2, 1.0, 1.0)            # :478
            contract("kbic,kjac->ijab", Wovov, t2, r2, -1.0, 1.0)           # :479
            contract("kaci,kjbc->ijab", Wovvo, t2, r2, -1.0, 1.0)           # :480
            contract("kbcj,ikac->ijab", Lovvo, t2, r2, 1.0, 1.0)            # :481
            contract("kajc,ikcb->ijab", Wovov, t2, r2, -1.0, 1.0)           # :482

        E = eng.run(residual, print_level, self.wfn.E_SCF, self.H.E_nuc)
        out = self._finish(eng, E, True)
        if config.VERBOSE and not config.RETURN_DEVICE:
            print("t-Amplitude Data:")
            print("Maximum t1: ", np.max(out[1]))
            print("Maximum t2: ", np.max(out[2]))
        return out
