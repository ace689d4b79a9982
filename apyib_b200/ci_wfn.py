"""CID / CISD wavefunctions -- drop-in for apyib/ci_wfn.py.

Same constructor (AO->MO transform and MO Fock build happen here, ci_wfn.py:46-47), same four
solve methods and return shapes.  Everything between "integrals are on the device" and
"converged amplitudes come back" runs in libapyib_b200 kernels:

  * every `oe.contract(...)` line of the reference's residuals is one launch of the DMMA
    contraction kernel (index regrouping via offset tables, no transposed copies);
  * `r -= E t; t += r/D`, the DIIS Gram row, the bordered solve, the extrapolation, the new energy
    and both rms sums are streaming / tiny kernels with deterministic reductions;
  * the iteration counter, energy and DIIS state live on the device, so one captured CUDA graph
    replays every iteration; the host only reads back 6 doubles per point and iteration for the
    convergence test (which keeps the reference's exact semantics, ci_wfn.py:123-130);
  * BATCHED SOLVES: `solve_batch` advances many finite-difference points (same shapes and
    dtype, each with its own integrals) with the same launches -- every tensor carries a leading
    point index that the contraction kernel treats as its batch dimension.  A point that meets
    the reference's convergence test is frozen by a device-side `active` flag, i.e. it stops
    exactly where the reference's loop `break`s.  The finite-difference driver uses this to fill the
    GPU with the 6N+7 tiny solves of one molecule.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import config
from ._lib import lib, check
from .contraction import contract, SCRATCH_OWNER
from .device import (to_device, to_host, empty, zeros, ptr, stream_ptr, Graph)
from .utils import (get_slices, compute_F_MO_dev, compute_ERI_MO_dev, spin_block_2_dev, gather4, gather4_stack,
                    mo_integrals_many)

_NULL = C.c_void_p(0)


def w_block_stack(Es, labels, out_labels, space, bounds, spin=0, c1=1.0, c2=0.0):
    """w_block for a list of tensors -> [len(Es), ...]; one launch when they are slices of one stack"""
    l0, l1, l2, l3 = labels
    src1, src2 = (l0, l2, l1, l3), (l0, l3, l1, l2)
    shape = [bounds[space[ch]][1] - bounds[space[ch]][0] for ch in out_labels]
    perm = lambda src: [out_labels.index(ch) for ch in src]
    start = lambda src: [bounds[space[ch]][0] for ch in src]
    return gather4_stack(Es, spin, shape, perm(src1), start(src1), c1, perm(src2), start(src2), c2)


def w_block(E, labels, out_labels, space, bounds, spin=0, c1=1.0, c2=0.0, out=None):
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
    return gather4(E, spin, shape, perm(src1), start(src1), c1, perm(src2), start(src2), c2, out=out)


class _Point:
    """MO integrals of one finite-difference point on the device (what ci_wfn.__init__ builds)."""

    def __init__(self, F, ERI, eps_o, eps_v, E_SCF=0.0, E_nuc=0.0, eps_dev=None):
        self.F, self.ERI, self.eps_o, self.eps_v, self.E_SCF, self.E_nuc = F, ERI, eps_o, eps_v, E_SCF, E_nuc
        self.eps_dev = eps_dev          # (occupied, virtual) orbital energies already on the device, or None


class _Engine:
    """Device-resident Jacobi/DIIS iteration shared by the four solvers, for nb points at once."""

    def __init__(self, parameters, dtype, nb, O, V, has_singles, spin_orbital, eps_o, eps_v, symmetrize=False):
        self.p = parameters
        self.dtype, self.nb, self.O, self.V = dtype, nb, O, V
        self.n1 = O * V if has_singles else 0
        self.n2 = O * O * V * V
        self.len = self.n1 + self.n2
        self.has_singles, self.so, self.symmetrize = has_singles, int(spin_orbital), symmetrize
        self.code = 1 if dtype == torch.complex128 else 0
        # [nb, o_spatial] / [nb, v_spatial] float64; device tensors when the points carry their uploaded energies
        up = lambda x: x if isinstance(x, torch.Tensor) else to_device(np.ascontiguousarray(np.real(x), dtype=np.float64))
        self.eps_o, self.eps_v = up(eps_o), up(eps_v)
        z = lambda *s: zeros(s, dtype)
        L = self.len
        self.t, self.r, self.t_old, self.w, self.r0 = z(nb, L), z(nb, L), z(nb, L), z(nb, L), z(nb, L)
        self.r_half = z(nb, self.n2) if symmetrize else None
        self.out6 = zeros((nb, 6), torch.float64)        # E(re,im), S1(re,im), S2(re,im) per point
        self.active = torch.ones((nb,), dtype=torch.int32, device=self.t.device)
        self.diis = bool(parameters["DIIS"])
        if self.diis:
            self.hist_e = z(nb, 8, L)
            self.hist_t = z(nb, 8, L)
            self.B = zeros((nb, 8 * 8 * 2), torch.float64)
            self.c = zeros((nb, 16), torch.float64)
            self.iter = torch.ones((1,), dtype=torch.int32, device=self.t.device)
        self.scratch = zeros((nb, int(lib.apyib_reduce_scratch_len())), torch.float64)
        self.iterations = [0] * nb

    # views into the concatenated vectors
    def t1(self, x=None):
        x = self.t if x is None else x
        return x[:, :self.n1].view(self.nb, self.O, self.V)

    def t2(self, x=None):
        x = self.t if x is None else x
        return x[:, self.n1:].view(self.nb, self.O, self.O, self.V, self.V)

    def _update(self):
        lr = getattr(self, "lr", None)      # linear-response mode: (E_fixed, E2_offset, t_fixed)
        if lr is None:
            check(lib.apyib_ci_update(self.code, ptr(self.r), ptr(self.t), ptr(self.out6), ptr(self.eps_o),
                                      ptr(self.eps_v), self.O, self.V, int(self.has_singles), self.so, self.nb,
                                      ptr(self.active), _NULL, _NULL, _NULL, stream_ptr()))
        else:
            E_fixed, E2_off, t_fixed = lr
            check(lib.apyib_ci_update(self.code, ptr(self.r), ptr(self.t), ptr(E_fixed), ptr(self.eps_o),
                                      ptr(self.eps_v), self.O, self.V, int(self.has_singles), self.so, self.nb,
                                      ptr(self.active), ptr(self.out6), ptr(E2_off), ptr(t_fixed), stream_ptr()))

    def _energy_rms(self, use_diis):
        if use_diis:
            check(lib.apyib_lincomb_energy_rms(self.code, ptr(self.hist_t), self.len, 0, ptr(self.iter), ptr(self.c),
                                               ptr(self.t), ptr(self.t_old), ptr(self.w), self.n1, self.len,
                                               ptr(self.out6), ptr(self.scratch), self.nb, ptr(self.active),
                                               stream_ptr()))
        else:
            check(lib.apyib_lincomb_energy_rms(self.code, _NULL, 0, 0, _NULL, _NULL, ptr(self.t), ptr(self.t_old),
                                               ptr(self.w), self.n1, self.len, ptr(self.out6), ptr(self.scratch),
                                               self.nb, ptr(self.active), stream_ptr()))

    def _copy(self, dst, src, alpha=1.0):
        """dst[s] = alpha * src[s] for equally shaped [nb, ...] tensors whose rows are dense: ONE launch for the batch"""
        if alpha == 1.0 and dst.is_contiguous() and src.is_contiguous():
            check(lib.apyib_copy(self.code, ptr(dst), ptr(src), dst.numel(), stream_ptr()))
        else:
            assert dst[0].is_contiguous() and src[0].is_contiguous()
            check(lib.apyib_copy_rows(self.code, ptr(dst), dst.stride(0), ptr(src), src.stride(0), dst[0].numel(),
                                      self.nb, float(alpha), _NULL, stream_ptr()))

    def initial_guess(self):
        """t = r0 / D, E = w.t   (ci_wfn.py:66-70, 193-197, 286-293, 435-442)."""
        if getattr(self, "custom_guess", None) is not None:
            self.custom_guess()
            return
        self.out6.zero_()
        self.t.zero_()
        self._copy(self.r, self.r0)
        if self.symmetrize:                   # CID spatial keeps 0.5*K in r0 (ci_wfn.py:83); guess uses K
            self._copy(self.r[:, self.n1:], self.r0[:, self.n1:], 2.0)
        self._update()
        self._copy(self.t_old, self.t)
        self._energy_rms(False)

    def iteration(self, residual):
        """One trip of the reference's while-loop body for all active points, on the current stream."""
        self._copy(self.t_old, self.t)
        if self.symmetrize:
            if self.n1:                                      # r1 <- its constant part; r2 is written by the symmetrisation
                self._copy(self.r[:, :self.n1], self.r0[:, :self.n1])
            self._copy(self.r_half, self.r0[:, self.n1:])
            residual(self.r_half)
            check(lib.apyib_symmetrize_ijab_batch(self.code, ptr(self.r_half), self.r_half.stride(0), ptr(self.r[:, self.n1:]),
                                                  self.r.stride(0), self.O, self.V, self.nb, ptr(self.active), stream_ptr()))
        else:
            self._copy(self.r, self.r0)
            residual(None)
        self._update()
        if self.diis:
            check(lib.apyib_diis_push(self.code, ptr(self.r), ptr(self.t), ptr(self.hist_e), ptr(self.hist_t),
                                      self.len, ptr(self.iter), ptr(self.B), ptr(self.scratch), self.nb,
                                      ptr(self.active), stream_ptr()))
            check(lib.apyib_diis_solve(self.code, ptr(self.B), 8, 0, ptr(self.iter), ptr(self.c), self.nb,
                                       ptr(self.active), stream_ptr()))
            self._energy_rms(True)
            check(lib.apyib_iter_advance(ptr(self.iter), stream_ptr()))
        else:
            self._energy_rms(False)

    def read(self):
        h = to_host(self.out6)
        if self.code:
            c = h[:, 0::2] + 1j * h[:, 1::2]
            return [tuple(np.complex128(x) for x in row) for row in c]
        return [tuple(np.float64(x) for x in row[0::2]) for row in h]

    def run(self, residual, print_level=0, E_SCF=0.0, E_nuc=0.0):
        g = self.run_steps(residual, print_level, E_SCF, E_nuc)
        while True:
            try:
                next(g)
            except StopIteration as done:
                return done.value

    def run_steps(self, residual, print_level=0, E_SCF=0.0, E_nuc=0.0):
        """The reference's iteration control (ci_wfn.py:76-130 etc.), verbatim semantics, applied to every point of
        the batch independently.  Generator: yields right after the launches of an iteration have been enqueued and
        BEFORE the blocking read-back of its 6 doubles per point, so that a caller can keep several batches (the
        float64 and the complex128 points of a molecule, each on its own stream) in flight from one host thread
        (`_drive`).  The generator's return value is the list of energies."""
        p, nb = self.p, self.nb
        self.initial_guess()
        yield
        E = [x[0] for x in self.read()]
        graph = None
        iteration = 1
        active = [True] * nb
        while iteration <= p["max_iterations"] and any(active):
            E_old = list(E)
            if iteration == 1 or not config.USE_CUDA_GRAPH:
                self.iteration(residual)              # eager: also builds/caches the offset tables
            else:
                if graph is None:
                    graph = Graph()
                    graph.capture(lambda: self.iteration(residual))
                graph.launch()
            yield
            vals = self.read()
            changed = False
            for s in range(nb):
                if not active[s]:
                    continue
                E[s], S1, S2 = vals[s]
                rms_t1, rms_t2 = np.sqrt(S1), np.sqrt(S2)
                delta_E = E_old[s] - E[s]
                self.iterations[s] = iteration
                if print_level > 0 and nb == 1:
                    E_tot = E_SCF + E[s] + E_nuc
                    print(" %02d %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f %20.12f" % (
                        iteration, np.real(E[s]), np.imag(E[s]), np.real(E_tot), np.real(delta_E), np.imag(delta_E),
                        np.real(rms_t1), np.imag(rms_t1), np.real(rms_t2), np.imag(rms_t2)))
                if iteration > 1:
                    conv = abs(delta_E) < p["e_convergence"] and rms_t2 < p["d_convergence"]
                    if self.has_singles:
                        conv = conv and rms_t1 < p["d_convergence"]
                    if conv:
                        active[s] = False
                        changed = True
                        continue
                if iteration == p["max_iterations"]:
                    if abs(delta_E) > p["e_convergence"] or rms_t2 > p["d_convergence"]:
                        print("Not converged.")
            if changed and any(active):
                self.active.copy_(torch.tensor([1 if a else 0 for a in active], dtype=torch.int32))
            iteration += 1
        return E


def _drive(jobs):
    """jobs: [(generator, stream | None)].  Advances the generators round-robin, each under its own stream, until all
    have finished; returns their return values.  While the host blocks on the read-back of one batch, the launches
    of the other batches are already queued on their streams and keep the GPU busy."""
    results = [None] * len(jobs)
    live = list(range(len(jobs)))
    while live:
        for k in list(live):
            gen, st = jobs[k]
            SCRATCH_OWNER[0] = None if st is None else k      # split-K scratch of this batch (contraction._work_buffer)
            try:
                if st is None:
                    next(gen)
                else:
                    with torch.cuda.stream(st):
                        next(gen)
            except StopIteration as done:
                results[k] = done.value
                live.remove(k)
            finally:
                SCRATCH_OWNER[0] = None
    return results


# -------------------------------------------------------------------------------------------------
# the four residuals, written for a stack of points (leading index s)
# -------------------------------------------------------------------------------------------------
def _sizes(points, so):
    n = points[0].F.shape[0]
    o = len(points[0].eps_o)
    f = 2 if so else 1
    O, V = f * o, f * (n - o)
    return O, V, {"o": (0, O), "v": (O, O + V)}


def _stack(items, fn, shape, dtype):
    out = empty((len(items),) + tuple(shape), dtype)
    for s, it in enumerate(items):
        fn(it, out[s])
    return out


def _make_engine(parameters, points, has_singles, so, symmetrize=False):
    O, V, bd = _sizes(points, so)
    if all(pt.eps_dev is not None for pt in points):          # no host->device copy on the solve path
        eps_o = torch.stack([pt.eps_dev[0] for pt in points])
        eps_v = torch.stack([pt.eps_dev[1] for pt in points])
    else:
        eps_o = np.stack([np.asarray(pt.eps_o) for pt in points])
        eps_v = np.stack([np.asarray(pt.eps_v) for pt in points])
    eng = _Engine(parameters, points[0].ERI.dtype, len(points), O, V, has_singles, so, eps_o, eps_v, symmetrize)
    return eng, O, V, bd


def _blocks(points, sp, bd, dt, spin):
    """returns blk(labels, out_labels=None, c1=1, c2=0): stacked [nb, ...] block of W (or <pq||rs>)"""
    def blk(lab, outl=None, c1=1.0, c2=0.0):
        return w_block_stack([pt.ERI for pt in points], lab, outl or lab, sp, bd, spin, c1, c2)
    return blk


def _solve_CID(parameters, points, print_level):
    """Spatial-orbital CID (ci_wfn.py:51-167)."""
    eng, O, V, bd = _make_engine(parameters, points, False, False, symmetrize=True)
    dt, nb = eng.dtype, eng.nb
    sp = dict(i="o", j="o", m="o", n="o", a="v", b="v", e="v", f="v")
    blk = _blocks(points, sp, bd, dt, 0)
    eng.r0.copy_(blk("abij", "ijab", 0.5).reshape(nb, -1))              # 0.5 <ab|ij>     ci_wfn.py:83
    eng.w.copy_(blk("ijab", None, 2.0, -1.0).reshape(nb, -1))           # 2<ij|ab>-<ij|ba> ci_wfn.py:70
    Woooo, Wvvvv = blk("mnij"), blk("abef")
    Wovvo, Wovov = blk("mbej", "mbje"), blk("mbie")                    # contracted e stored last (see _CISDOperator)
    Lovvo = blk("mbej", "mbje", 1.0, -1.0)                               # <mb|ej>-<mb|je>  ci_wfn.py:89
    F = torch.stack([pt.F for pt in points])
    Foo, Fvv = F[:, :O, :O], F[:, O:, O:]

    ladder = _PackedLadder(Wvvvv, O, V, nb, dt) if config.PACKED_LADDER else None

    def residual(r):
        t2 = eng.t2()
        r = r.view(nb, O, O, V, V)
        contract("sijae,sbe->sijab", t2, Fvv, r, 1.0, 1.0)              # ci_wfn.py:84
        contract("simab,smj->sijab", t2, Foo, r, -1.0, 1.0)             # :85
        contract("smnab,smnij->sijab", t2, Woooo, r, 0.5, 1.0)          # :86
        if ladder is not None:                                          # :87 over the pairs i <= j only
            ladder.apply(t2, r)
        else:
            contract("sijef,sabef->sijab", t2, Wvvvv, r, 0.5, 1.0)
        contract("simae,smbje->sijab", t2, Wovvo, r, 1.0, 1.0)          # :88  (t2 - t2.swapaxes(2,3)) . W
        contract("simea,smbje->sijab", t2, Wovvo, r, -1.0, 1.0)
        contract("simae,smbje->sijab", t2, Lovvo, r, 1.0, 1.0)          # :89
        contract("smjae,smbie->sijab", t2, Wovov, r, -1.0, 1.0)         # :90

    E = yield from eng.run_steps(residual, print_level, points[0].E_SCF, points[0].E_nuc)
    return eng, E


def _solve_CID_SO(parameters, points, print_level):
    """Spin-orbital CID (ci_wfn.py:171-259)."""
    eng, O, V, bd = _make_engine(parameters, points, False, True)
    dt, nb = eng.dtype, eng.nb
    sp = dict(i="o", j="o", m="o", n="o", a="v", b="v", e="v", f="v")
    blk = _blocks(points, sp, bd, dt, 1)
    A = lambda lab, outl=None, sc=1.0: blk(lab, outl, sc, -sc)            # <pq||rs>
    F = torch.stack([spin_block_2_dev(pt.F) for pt in points])             # compute_F_SO, ci_wfn.py:185
    eng.r0.copy_(A("abij", "ijab").reshape(nb, -1))                       # ci_wfn.py:210
    eng.w.copy_(A("ijab", None, 0.25).reshape(nb, -1))                    # ci_wfn.py:197
    Aoooo, Avvvv, Aovvo = A("mnij"), A("abef"), A("mbej", "mbje")
    Foo, Fvv = F[:, :O, :O], F[:, O:, O:]

    def residual(_):
        t2, r = eng.t2(), eng.t2(eng.r)
        contract("sijae,sbe->sijab", t2, Fvv, r, 1.0, 1.0)                # :211
        contract("sijeb,sae->sijab", t2, Fvv, r, 1.0, 1.0)
        contract("simab,smj->sijab", t2, Foo, r, -1.0, 1.0)               # :212
        contract("smjab,smi->sijab", t2, Foo, r, -1.0, 1.0)
        contract("smnab,smnij->sijab", t2, Aoooo, r, 0.5, 1.0)            # :213
        contract("sijef,sabef->sijab", t2, Avvvv, r, 0.5, 1.0)            # :214
        contract("simae,smbje->sijab", t2, Aovvo, r, 1.0, 1.0)            # :215
        contract("smjae,smbie->sijab", t2, Aovvo, r, 1.0, 1.0)            # :216
        contract("simeb,smaje->sijab", t2, Aovvo, r, 1.0, 1.0)            # :217
        contract("smjeb,smaie->sijab", t2, Aovvo, r, 1.0, 1.0)            # :218

    E = yield from eng.run_steps(residual, print_level, points[0].E_SCF, points[0].E_nuc)
    return eng, E


def _solve_CISD_SO(parameters, points, print_level):
    """Spin-orbital CISD (ci_wfn.py:263-416)."""
    eng, O, V, bd = _make_engine(parameters, points, True, True)
    dt, nb, n1 = eng.dtype, eng.nb, eng.n1
    sp = dict(i="o", j="o", k="o", l="o", a="v", b="v", c="v", d="v")
    blk = _blocks(points, sp, bd, dt, 1)
    A = lambda lab, outl=None, sc=1.0: blk(lab, outl, sc, -sc)
    F = torch.stack([spin_block_2_dev(pt.F) for pt in points])
    Foo, Fvv, Fov = F[:, :O, :O], F[:, O:, O:], F[:, :O, O:].contiguous()
    eng.r0[:, :n1].copy_(F[:, O:, :O].transpose(1, 2).reshape(nb, -1))     # F_ai as [i,a]   ci_wfn.py:309
    eng.r0[:, n1:].copy_(A("abij", "ijab").reshape(nb, -1))               # :319
    eng.w[:, :n1].copy_(Fov.reshape(nb, -1))                              # :355
    eng.w[:, n1:].copy_(A("ijab", None, 0.25).reshape(nb, -1))
    Aovvo, Avovv, Aooov = A("jabi", "jaib"), A("ajcb"), A("kjib")       # layouts: see _CISDOperator
    Aovoo, Avooo, Avvvo, Avvov = A("kbij", "kijb"), A("akij", "kija"), A("abcj", "jabc"), A("abic", "iabc")
    Aoooo, Avvvv = A("klij"), A("abcd")

    def residual(_):
        t1, t2 = eng.t1(), eng.t2()
        r1, r2 = eng.t1(eng.r), eng.t2(eng.r)
        contract("sji,sja->sia", Foo, t1, r1, -1.0, 1.0)                  # :310
        contract("sab,sib->sia", Fvv, t1, r1, 1.0, 1.0)                   # :311
        contract("sjaib,sjb->sia", Aovvo, t1, r1, 1.0, 1.0)               # :312
        contract("sjb,sijab->sia", Fov, t2, r1, 1.0, 1.0)                 # :313
        contract("sajcb,sijcb->sia", Avovv, t2, r1, 0.5, 1.0)             # :314
        contract("skjib,skjab->sia", Aooov, t2, r1, -0.5, 1.0)            # :315
        contract("skijb,ska->sijab", Aovoo, t1, r2, -1.0, 1.0)            # :320
        contract("skija,skb->sijab", Avooo, t1, r2, -1.0, 1.0)            # :321
        contract("sjabc,sic->sijab", Avvvo, t1, r2, 1.0, 1.0)             # :322
        contract("siabc,sjc->sijab", Avvov, t1, r2, 1.0, 1.0)             # :323
        contract("sbc,sijac->sijab", Fvv, t2, r2, 1.0, 1.0)               # :324
        contract("sac,sijcb->sijab", Fvv, t2, r2, 1.0, 1.0)               # :325
        contract("skj,sikab->sijab", Foo, t2, r2, -1.0, 1.0)              # :326
        contract("ski,skjab->sijab", Foo, t2, r2, -1.0, 1.0)              # :327
        contract("sklij,sklab->sijab", Aoooo, t2, r2, 0.5, 1.0)           # :328
        contract("sabcd,sijcd->sijab", Avvvv, t2, r2, 0.5, 1.0)           # :329
        contract("skbjc,sikac->sijab", Aovvo, t2, r2, 1.0, 1.0)           # :330
        contract("skbic,skjac->sijab", Aovvo, t2, r2, 1.0, 1.0)           # :331
        contract("skajc,sikcb->sijab", Aovvo, t2, r2, 1.0, 1.0)           # :332
        contract("skaic,skjcb->sijab", Aovvo, t2, r2, 1.0, 1.0)           # :333

    E = yield from eng.run_steps(residual, print_level, points[0].E_SCF, points[0].E_nuc)
    return eng, E


class _PackedLadder:
    """The (i<->j, a<->b)-symmetric ladder term  L_ijab = sum_cd <ab|cd> t_ijcd  (ci_wfn.py:87, 476) for the half-sum
    residual form r2 = h + P h: instead of adding L/2 for all o^2 occupied pairs, add L for i < j and L/2 for i == j
    -- o(o+1)/2 pairs, 46 % fewer flops at o = 12 -- and let the symmetrisation supply the rest
    ((h + P h)_jiba = L_ijab = L_jiba).  t2 is packed over the pairs (diagonal pre-scaled by 1/2), contracted with
    the TMA-fed DMMA kernel and scatter-added into h."""

    def __init__(self, W, O, V, nb, dt):
        self.W, self.O, self.V, self.nb = W, O, V, nb
        self.npair = O * (O + 1) // 2
        self.code = 1 if dt == torch.complex128 else 0
        self.tp = empty((nb, self.npair, V, V), dt)
        self.hp = empty((nb, self.npair, V, V), dt)

    def apply(self, t2, h):
        """h[s,i,j,a,b] += w_ij L[s,i,j,a,b] for i <= j (t2, h: [nb, O, O, V, V] views with dense (i,j,a,b) blocks)"""
        vv = self.V * self.V
        check(lib.apyib_pack_pairs(self.code, ptr(t2), t2.stride(0), ptr(self.tp), self.tp.stride(0), self.O, vv,
                                   self.nb, _NULL, stream_ptr()))
        contract("sabcd,spcd->spab", self.W, self.tp, self.hp, 1.0, 0.0)
        check(lib.apyib_unpack_pairs_add(self.code, ptr(self.hp), self.hp.stride(0), ptr(h), h.stride(0), self.O, vv,
                                         self.nb, _NULL, stream_ptr()))


class _CISDOperator:
    """Spatial-orbital CISD residual of ci_wfn.py:457-483 for a stack of points, split into its
    constant part (F_ai | <ab|ij>), its energy weights (2 F_ov | 2<ij|ab>-<ij|ba>) and the part that is
    linear in the amplitudes -- the same contraction set serves the CISD solver and the
    perturbed-amplitude (linear-response) iterations of analytic_aats.py:780-885 / 1032-1137."""

    def __init__(self, Fs, ERIs, O, V, bd, dt):
        nb = len(ERIs)
        self.nb, self.O, self.V, self.n1 = nb, O, V, O * V
        sp = dict(i="o", j="o", k="o", l="o", a="v", b="v", c="v", d="v")
        blk = lambda lab, outl=None, c1=1.0, c2=0.0: w_block_stack(list(ERIs), lab, outl or lab, sp, bd, 0, c1, c2)
        F = torch.stack(list(Fs))
        self.Foo, self.Fvv, self.Fov = F[:, :O, :O], F[:, O:, O:], F[:, :O, O:].contiguous()
        self.Fai = F[:, O:, :O].transpose(1, 2).reshape(nb, -1)                 # F_ai as [i,a]  ci_wfn.py:457
        self.K = blk("abij", "ijab").reshape(nb, -1)                             # <ab|ij>        :466
        w1 = empty((nb, self.n1), dt)
        check(lib.apyib_axpby(1 if dt == torch.complex128 else 0, nb * self.n1, 2.0, 0.0, ptr(self.Fov), 0, 0.0, 0.0,
                              ptr(w1), stream_ptr()))                            # 2 F_ov         :504
        self.w1 = w1
        self.w2 = blk("ijab", None, 2.0, -1.0).reshape(nb, -1)                   # 2<ij|ab>-<ij|ba>
        # Block layouts are chosen for the contraction kernel, not copied from the reference's slices: the
        # contracted virtual index is stored LAST (contiguous, 70-element runs at cc-pVDZ sizes) and the
        # free indices in the order they have in r2[i,j,a,b], so every A-operand gather is coalesced (a
        # 12-element run of an occupied index costs a full 32-byte sector per 8-byte element otherwise:
        # 19.7 -> 25.7 TFLOP/s on the ring terms, 1.2 -> 4 TB/s on the T1 couplings).
        self.Wovvo, self.Wovov = blk("kbcj", "kbjc"), blk("kbic")
        self.Lovvo = blk("jabi", "jaib", 2.0, -1.0)                              # 2<ja|bi>-<ja|ib>  :460,478,481
        self.Lvovv = blk("ajbc", None, 2.0, -1.0)                                # :462
        self.Looov = blk("kjib", None, 2.0, -1.0)                                # :463
        self.Wvvvo, self.Wvvov = blk("abcj", "jabc"), blk("abic", "iabc")
        self.Wovoo, self.Wvooo = blk("kbij", "kijb"), blk("akij", "kija")
        self.Woooo, self.Wvvvv = blk("klij"), blk("abcd")
        self._ladder = None

    def apply(self, t1, t2, r1, r2, half=False, singles=True):
        """r += (linear part of the CISD residual)(t1, t2).  singles=False keeps only the doubles <- doubles
        terms (t1, r1 unused): the CID residual in the 12-term form of analytic_aats.py:1588-1600.

        half=True: r2 receives only h with (linear part) = h + P h, P = (i<->j, a<->b): the reference's 16
        r_T2 terms (ci_wfn.py:467-482) are the ladder and the oooo term, which are P-symmetric, and seven
        pairs (X, P X) -- :467/:468, :469/:470, :471/:472, :473/:474, :477/:480, :478/:481, :479/:482 -- by
        <pq|rs> = <qp|sr> and t_ijab = t_jiba.  This is the half-sum-then-symmetrise form the reference
        itself uses for CID (ci_wfn.py:83-92); the caller applies r2 = h + P h (apyib_symmetrize_ijab).
        It halves the ring and coupling work of every iteration."""
        o = self
        c = 0.5 if half else 1.0
        if singles:
            contract("sji,sja->sia", o.Foo, t1, r1, -1.0, 1.0)              # :458
            contract("sab,sib->sia", o.Fvv, t1, r1, 1.0, 1.0)               # :459
            contract("sjaib,sjb->sia", o.Lovvo, t1, r1, 1.0, 1.0)           # :460
            contract("sjb,sijab->sia", o.Fov, t2, r1, 2.0, 1.0)             # :461  F.(2 t2 - t2^T)
            contract("sjb,sijba->sia", o.Fov, t2, r1, -1.0, 1.0)
            contract("sajbc,sijbc->sia", o.Lvovv, t2, r1, 1.0, 1.0)         # :462
            contract("skjib,skjab->sia", o.Looov, t2, r1, -1.0, 1.0)        # :463
            contract("sjabc,sic->sijab", o.Wvvvo, t1, r2, 1.0, 1.0)         # :467
            contract("skijb,ska->sijab", o.Wovoo, t1, r2, -1.0, 1.0)        # :469
        contract("sac,sijcb->sijab", o.Fvv, t2, r2, 1.0, 1.0)               # :471
        contract("ski,skjab->sijab", o.Foo, t2, r2, -1.0, 1.0)              # :473
        contract("sklij,sklab->sijab", o.Woooo, t2, r2, c, 1.0)             # :475
        if half and config.PACKED_LADDER:                                   # :476 over the pairs i <= j only
            if o._ladder is None:
                o._ladder = _PackedLadder(o.Wvvvv, o.O, o.V, o.nb, t2.dtype)
            o._ladder.apply(t2, r2)
        else:
            contract("sabcd,sijcd->sijab", o.Wvvvv, t2, r2, c, 1.0)         # :476
        contract("skbjc,sikca->sijab", o.Wovvo, t2, r2, -1.0, 1.0)          # :477
        contract("skaic,skjcb->sijab", o.Lovvo, t2, r2, 1.0, 1.0)           # :478
        contract("skbic,skjac->sijab", o.Wovov, t2, r2, -1.0, 1.0)          # :479
        if half:
            return
        if singles:
            contract("siabc,sjc->sijab", o.Wvvov, t1, r2, 1.0, 1.0)         # :468
            contract("skija,skb->sijab", o.Wvooo, t1, r2, -1.0, 1.0)        # :470
        contract("sbc,sijac->sijab", o.Fvv, t2, r2, 1.0, 1.0)               # :472
        contract("skj,sikab->sijab", o.Foo, t2, r2, -1.0, 1.0)              # :474
        contract("skaic,skjbc->sijab", o.Wovvo, t2, r2, -1.0, 1.0)          # :480
        contract("skbjc,sikac->sijab", o.Lovvo, t2, r2, 1.0, 1.0)           # :481
        contract("skajc,sikcb->sijab", o.Wovov, t2, r2, -1.0, 1.0)          # :482


def _solve_CISD(parameters, points, print_level):
    """Spatial-orbital CISD (ci_wfn.py:420-574).  r_T2 is built as h + P h (see _CISDOperator.apply)."""
    eng, O, V, bd = _make_engine(parameters, points, True, False, symmetrize=True)
    n1, nb = eng.n1, eng.nb
    op = _CISDOperator([pt.F for pt in points], [pt.ERI for pt in points], O, V, bd, eng.dtype)
    eng.r0[:, :n1].copy_(op.Fai)
    eng._copy(eng.r0[:, n1:], op.K, 0.5)      # the engine keeps K/2 and symmetrises (ci_wfn.py:83 form)
    eng.w[:, :n1].copy_(op.w1)
    eng.w[:, n1:].copy_(op.w2)
    residual = lambda rh: op.apply(eng.t1(), eng.t2(), eng.t1(eng.r), rh.view(nb, O, O, V, V), half=True)
    E = yield from eng.run_steps(residual, print_level, points[0].E_SCF, points[0].E_nuc)
    return eng, E


def solve_perturbed_CISD(parameters, ci, t1, t2, E_CISD, dF_MO, dERI_MO, dE_guess=0.0, print_level=0):
    """Perturbed-amplitude (linear-response) iterations of the analytic CISD AAT route,
    analytic_aats.py:742-885 (magnetic field) and :994-1137 (nuclear displacement): given the
    converged spatial CISD wavefunction (t1, t2, E_CISD) of `ci` and the perturbed MO Fock matrix /
    chemists' MO integrals (dF_MO, dERI_MO: host inputs from CPHF + Psi4 derivative integrals),
    solve  dR(t; dF, dERI) - dE t + R_lin(dt; F, ERI) - E_CISD dt = 0  for (dt1, dt2) with the
    reference's Jacobi/DIIS iteration.  `dE_guess` is the relaxed energy derivative the reference
    uses for its starting guess (:743, :759).  Returns (dE_proj, dt1, dt2).

    The part of the residual that only involves the unperturbed amplitudes is built once (the
    reference rebuilds it every iteration); each iteration then costs one CISD contraction set."""
    return _solve_perturbed(parameters, ci, t1, t2, E_CISD, dF_MO, dERI_MO, dE_guess, print_level, True)


def solve_perturbed_CID(parameters, ci, t2, E_CID, dF_MO, dERI_MO, dE_guess=0.0, print_level=0):
    """The CID variant, analytic_aats.py:1553-1649 (magnetic field) and :1754-1850 (nuclear
    displacement): the same iteration restricted to the doubles <- doubles terms (:1588-1614),
    convergence on the energy derivative and rms(dt2) only (:1636-1639).  Returns (dE_proj, dt2)."""
    dE, _, dt2 = _solve_perturbed(parameters, ci, None, t2, E_CID, dF_MO, dERI_MO, dE_guess, print_level, False)
    return dE, dt2


def _solve_perturbed(parameters, ci, t1, t2, E_CI, dF_MO, dERI_MO, dE_guess, print_level, singles):
    pt = ci.point()
    dt_ = pt.ERI.dtype
    cast = lambda x: to_device(np.asarray(x), dt_)
    dF, dERI = cast(dF_MO), cast(dERI_MO)
    eng, O, V, bd = _make_engine(parameters, [pt], singles, False)
    n1, L = eng.n1, eng.len
    op0 = _CISDOperator([pt.F], [pt.ERI], O, V, bd, dt_)
    op1 = _CISDOperator([dF], [dERI], O, V, bd, dt_)
    tfix = zeros((1, L), dt_)
    if singles:
        tfix[:, :n1].copy_(cast(t1).reshape(1, -1))
    tfix[:, n1:].copy_(cast(t2).reshape(1, -1))
    # constant part: dF_ai | d<ab|ij>  +  (perturbed integrals) x (unperturbed amplitudes)
    t1v = (lambda x=None: eng.t1(x)) if singles else (lambda x=None: None)
    if singles:
        eng.r0[:, :n1].copy_(op1.Fai)
        eng.w[:, :n1].copy_(op0.w1)
    eng.r0[:, n1:].copy_(op1.K)
    op1.apply(t1v(tfix), eng.t2(tfix), t1v(eng.r0), eng.t2(eng.r0), singles=singles)
    eng.w[:, n1:].copy_(op0.w2)
    # constant part of the projected energy derivative: 2 t1.dF_ov + t2.(2 d<ij|ab> - d<ij|ba>)  (:774, :868)
    wd = zeros((1, L), dt_)
    if singles:
        wd[:, :n1].copy_(op1.w1)
    wd[:, n1:].copy_(op1.w2)
    c0 = zeros((2,), torch.float64)
    check(lib.apyib_dots(eng.code, ptr(wd), 0, 1, ptr(tfix), L, 0, ptr(c0), ptr(eng.scratch), stream_ptr()))
    E2_off = zeros((1, 6), torch.float64)
    E2_off[0, :2].copy_(c0)
    E_fixed = zeros((1, 6), torch.float64)
    Ec = complex(E_CI)
    E_fixed[0, 0], E_fixed[0, 1] = Ec.real, Ec.imag
    c0h = to_host(c0)
    c0v = complex(c0h[0], c0h[1]) if eng.code else float(c0h[0])

    def guess():
        # dt = [ -dE_guess t + (perturbed integrals) x t ] / D      (:743-772; no dF_ai / d<ab|ij> term)
        eng.out6.zero_()
        eng.t.zero_()
        eng._copy(eng.r, eng.r0)
        g = complex(dE_guess)
        a = lambda alpha, x, y: check(lib.apyib_axpby(eng.code, x.numel(), alpha.real, alpha.imag, ptr(x), 0, 1.0, 0.0,
                                                     ptr(y), stream_ptr()))
        cons = zeros((1, L), dt_)
        if singles:
            cons[:, :n1].copy_(op1.Fai)
        cons[:, n1:].copy_(op1.K)
        a(complex(-1.0), cons, eng.r)
        a(-g, tfix, eng.r)
        eng.lr = None
        eng._update()                                  # E = 0: dt = r / D
        eng.lr = (E_fixed, E2_off, tfix)
        eng._copy(eng.t_old, eng.t)
        eng._energy_rms(False)

    eng.custom_guess = guess
    eng.lr = (E_fixed, E2_off, tfix)
    residual = lambda _: op0.apply(t1v(), eng.t2(), t1v(eng.r), eng.t2(eng.r), singles=singles)
    E = eng.run(residual, print_level)
    dE = E[0] + c0v
    ci.iterations = eng.iterations[0]
    if config.RETURN_DEVICE:
        return dE, (eng.t1()[0].clone() if singles else None), eng.t2()[0].clone()
    return dE, (to_host(eng.t1())[0].copy() if singles else None), to_host(eng.t2())[0].copy()


_SOLVERS = {"CID": (_solve_CID, False), "CID_SO": (_solve_CID_SO, False),
            "CISD": (_solve_CISD, True), "CISD_SO": (_solve_CISD_SO, True)}


def _collect(eng, E, singles):
    """per-point result tuples (E, t2) / (E, t1, t2) of a finished engine, host or device-resident"""
    res = []
    dev_out = config.RETURN_DEVICE
    t1s = eng.t1() if singles else None
    t2s = eng.t2()
    if not dev_out:
        t2h = to_host(t2s)
        t1h = to_host(t1s) if singles else None
    for s in range(eng.nb):
        if dev_out:
            t2 = t2s[s].clone()
            res.append((E[s], t1s[s].clone(), t2) if singles else (E[s], t2))
        else:           # disjoint views of the batch's (pinned) host block: no second host copy
            res.append((E[s], t1h[s], t2h[s]) if singles else (E[s], t2h[s]))
    return res


def solve_batch(method, parameters, points, print_level=0):
    """Solve `method` for a list of _Point objects (identical shapes and dtype) with shared
    launches.  Returns (results, iterations): results[s] = (E, t2) or (E, t1, t2) exactly as the
    single-point methods return them."""
    fn, singles = _SOLVERS[method]
    eng, E = _drive([(fn(parameters, points, print_level), None)])[0]
    return _collect(eng, E, singles), list(eng.iterations)


_side_streams = {}
_DEBUG_KEEP = None          # tools/diag_race3.py: list that receives the ci_wfn objects of every solve_many call


def _streams(n):
    """n side streams of the current device, created once: PyTorch's caching allocator keeps one memory pool per
    stream, so fresh streams per solve would mean fresh cudaMallocs per solve."""
    d = torch.cuda.current_device()
    pool = _side_streams.setdefault(d, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=d))
    return pool[:n]


def _concurrent(sizes):
    """config.SOLVE_CONCURRENT -> run these groups (sizes = elements per group: points x o^2 v^2) concurrently?"""
    mode = config.SOLVE_CONCURRENT
    if mode in (True, "1", 1):
        return True
    if mode in (False, "0", 0, None):
        return False
    return len(sizes) > 1 and max(sizes) <= config.SOLVE_CONCURRENT_MAX_ELEMS


def _run_jobs(make_jobs, singles, concurrent=False):
    """make_jobs: list of callables, each returning the generator of one batch (its set-up runs at the generator's
    first step, under the batch's stream).  Sequential: one batch after the other on the caller's stream.
    Concurrent: one side stream per batch, driven from this host thread, with the TMA-fed contraction kernel
    switched off for the duration (config.SOLVE_CONCURRENT).  Returns [(results, iterations)] per batch."""
    if len(make_jobs) == 1 or not concurrent:
        out = []
        for mk in make_jobs:
            eng, E = _drive([(mk(), None)])[0]
            out.append((_collect(eng, E, singles), list(eng.iterations)))
        return out
    use_tma, config.USE_TMA = config.USE_TMA, False
    try:
        return _run_jobs_concurrently(make_jobs, singles)
    finally:
        config.USE_TMA = use_tma


def _run_jobs_concurrently(make_jobs, singles):
    cur = torch.cuda.current_stream()
    streams = _streams(len(make_jobs))
    for st in streams:
        st.wait_stream(cur)                    # inputs may have been produced on the caller's stream
    done = _drive([(mk(), st) for mk, st in zip(make_jobs, streams)])
    out = []
    for (eng, E), st in zip(done, streams):
        with torch.cuda.stream(st):
            res = _collect(eng, E, singles)
        cur.wait_stream(st)
        for r in res:                          # device-resident results were allocated on a side stream
            for x in r:
                if isinstance(x, torch.Tensor) and x.is_cuda:
                    x.record_stream(cur)
        out.append((res, list(eng.iterations)))
    return out


def solve_batches(method, parameters, batches, print_level=0):
    """Several independent batches (e.g. the float64 and the complex128 finite-difference points of one molecule)
    solved CONCURRENTLY: one CUDA stream per batch, all driven from this host thread (`_drive`) -- every batch is
    latency bound on its own (a chain of ~40 dependent launches and one 48-byte read-back per iteration), together
    they fill the device.  Results are identical to solving the batches one after the other (same launches, same
    order within a batch).  Returns [(results, iterations)] per batch."""
    fn, singles = _SOLVERS[method]
    return _run_jobs([(lambda pts=pts: fn(parameters, pts, print_level)) for pts in batches], singles,
                     concurrent=config.SOLVE_CONCURRENT in (True, "1", 1))


class ci_wfn(object):
    """Reference: apyib/ci_wfn.py:19-575."""

    def __init__(self, parameters, wfn, _integrals=None):
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
        if _integrals is None:
            self._F_dev, self._E_fc = compute_F_MO_dev(self.parameters, self.wfn, self.C_list)  # ci_wfn.py:46
            self._ERI_dev = compute_ERI_MO_dev(self.parameters, self.wfn, self.C_list)         # ci_wfn.py:47
        else:                               # built for a whole stack of points at once (ci_wfn.many)
            self._F_dev, self._E_fc, self._ERI_dev = _integrals
        self._F_host = self._ERI_host = None
        self.iterations = 0

    @classmethod
    def many(cls, parameters, wfns):
        """[ci_wfn(parameters, w) for w in wfns] with the AO->MO transforms and Fock builds of all points of one
        shape and dtype done by shared launches (utils.mo_integrals_many)."""
        C_lists = [get_slices(parameters, w)[0] for w in wfns]
        ints = mo_integrals_many(parameters, wfns, C_lists)
        return [cls(parameters, w, _integrals=i) for w, i in zip(wfns, ints)]

    @property
    def E_fc(self):
        """frozen-core energy of utils.compute_F_MO (utils.py:238-243); read back from the device on first use"""
        if hasattr(self._E_fc, "get"):
            self._E_fc = self._E_fc.get()
        return self._E_fc

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

    def point(self):
        """Device-side integrals of this wavefunction, for solve_batch."""
        from .utils import wfn_small_on_device
        eps = wfn_small_on_device(self.wfn, self._ERI_dev.dtype == torch.complex128)[1]
        o, v = self.C_list[1], self.C_list[2]
        return _Point(self._F_dev, self._ERI_dev, self.eps_o, self.eps_v, self.wfn.E_SCF, self.H.E_nuc,
                      eps_dev=(eps[o.start:o.stop], eps[v.start:v.stop]))

    def _solve(self, method, print_level):
        res, its = solve_batch(method, self.parameters, [self.point()], print_level)
        self.iterations = its[0]
        return res[0]

    def solve_CID(self, print_level=0):
        """Spatial-orbital CID (ci_wfn.py:51-167).  Returns (E_CID, t2)."""
        out = self._solve("CID", print_level)
        if config.VERBOSE and not config.RETURN_DEVICE:
            print("t-Amplitude Data:")
            print("Maximum t2: ", np.max(out[1]))
        return out

    def solve_CID_SO(self, print_level=0):
        """Spin-orbital CID (ci_wfn.py:171-259).  Returns (E_CID, t2) in the spin-orbital basis."""
        return self._solve("CID_SO", print_level)

    def solve_CISD_SO(self, print_level=0):
        """Spin-orbital CISD (ci_wfn.py:263-416).  Returns (E_CISD, t1, t2)."""
        return self._solve("CISD_SO", print_level)

    def solve_CISD(self, print_level=0):
        """Spatial-orbital CISD (ci_wfn.py:420-574).  Returns (E_CISD, t1, t2)."""
        out = self._solve("CISD", print_level)
        if config.VERBOSE and not config.RETURN_DEVICE:
            print("t-Amplitude Data:")
            print("Maximum t1: ", np.max(out[1]))
            print("Maximum t2: ", np.max(out[2]))
        return out


def solve_many(method, parameters, wfns, print_level=0):
    """Batched counterpart of `[ci_wfn(parameters, w).solve_<method>() for w in wfns]`.  The points are grouped by
    dtype (real nuclear-displacement points / complex field points); the host->device copies of their AO integrals
    are queued on the copy stream up front (complex points first), and every group runs as its own pipeline -- MO
    integrals (each point's transform waits for its own upload), integral blocks, iterations -- on its own stream,
    all groups driven concurrently from this thread: the complex group is already iterating while the real points
    are still uploading.  Returns the per-point result tuples in input order."""
    from .utils import ao_prefetch, _is_complex
    fn, singles = _SOLVERS[method]
    groups = {}
    for k, w in enumerate(wfns):
        nf = w.H.basis_set.n_frozen_core()
        groups.setdefault((0 if _is_complex(w) else 1, int(w.nbf), int(w.ndocc), nf), []).append(k)
    keys = sorted(groups)                                            # complex groups first
    ao_prefetch([wfns[k] for key in keys for k in groups[key]])
    idxs = []
    for key in keys:                                                 # big shapes: chunks of one dtype, in upload order
        idx = groups[key]
        big = 8 * key[1] ** 4 >= config.SOLVE_CHUNK_MIN_BYTES and len(idx) > config.SOLVE_CHUNK
        nchunk = -(-len(idx) // config.SOLVE_CHUNK) if big else 1
        size = -(-len(idx) // nchunk)
        idxs += [idx[i:i + size] for i in range(0, len(idx), size)]
    cis = [None] * len(wfns)
    if _DEBUG_KEEP is not None:
        _DEBUG_KEEP.append(cis)

    def job(idx):
        for k, c in zip(idx, ci_wfn.many(parameters, [wfns[k] for k in idx])):
            cis[k] = c
        return (yield from fn(parameters, [cis[k].point() for k in idx], print_level))

    def elems(idx):                                                  # points x o^2 v^2 of a group
        w = wfns[idx[0]]
        o = int(w.ndocc) - w.H.basis_set.n_frozen_core()
        v = int(w.nbf) - int(w.ndocc)
        f = 4 if method.endswith("_SO") else 1
        return len(idx) * f * f * o * o * v * v

    solved = _run_jobs([(lambda idx=idx: job(idx)) for idx in idxs], singles,
                       concurrent=_concurrent([elems(idx) for idx in idxs]))
    out = [None] * len(wfns)
    for idx, (res, its) in zip(idxs, solved):
        for k, r, it in zip(idx, res, its):
            out[k] = r
            cis[k].iterations = it
    return out
