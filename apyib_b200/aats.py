"""Finite-difference atomic axial tensors -- drop-in for apyib/aats.py.

Public surface kept: AAT(...) constructor, compute_SO_det, compute_normalization,
compute_all_dets, compute_spatial_aats, compute_SO_aats (+ the nine compute_SO_I_* terms).

B200 design:
  * every substituted determinant is an LU done by a sub-warp on the device
    (csrc/dets.cu); substituted matrices are formed on the fly from the MO overlap and index lists;
  * the reference materialises the antisymmetrised determinant tensors (8 indices,
    aats.py:575/630) and then contracts them with amplitudes.  Here the antisymmetric completion
    is folded into the amplitude vectors (apyib_pack_doubles) and the det-table x vector
    products are fused into the LU kernel (apyib_det_matvec) -- the tables never exist;
  * the remaining o^2 v^2-sized contractions go through the DMMA contraction kernel;
  * determinant families that do not depend on (alpha, beta) -- uu, up/un[beta], pu/nu[alpha] -- are
    evaluated once per molecule for all bra/ket amplitude vectors and cached, instead of being
    recomputed for every tensor element as `compute_spatial_aats` does in the reference.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import config
from ._lib import lib, check, LAUNCHES as _liblaunch
from .contraction import contract, contract_new
from .device import to_device, to_host, empty, zeros, ptr, stream_ptr, reduce_scratch, device
from .utils import mo_overlaps_dev, spin_block_2_dev, SO_METHODS

_C128 = torch.complex128
_tables = {}


def _i32_host(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def prefix_groups(srt, n, k):
    """Structure the prefix-shared LU kernel (csrc/dets_pairs.cu) relies on, verified on the host: the sorted
    lists come in groups of consecutive lists sharing their first n-k columns and ending in every candidate
    column (k = 1) / every pair c < d in lexicographic order (k = 2) of one ascending candidate set.
    Returns (group length, candidate columns int32[nc], nc) or None when the lists are not of that form."""
    if n <= k or len(srt) == 0:
        return None
    cand = np.unique(srt[:, n - k:])
    nc = len(cand)
    if k == 1:
        tails = cand.reshape(-1, 1)
    else:
        if nc < 2:
            return None
        tails = np.array([(cand[x], cand[y]) for x in range(nc) for y in range(x + 1, nc)], dtype=srt.dtype)
    gl = len(tails)
    if len(srt) % gl:
        return None
    grp = srt.reshape(-1, gl, n)
    if not (np.array_equal(grp[:, :, :n - k], np.repeat(grp[:, :1, :n - k], gl, axis=1))
            and np.array_equal(grp[:, :, n - k:], np.broadcast_to(tails, (grp.shape[0], gl, k)))
            and (grp[:, :, :n - k] < n).all() and (cand >= n).all()):
        return None
    return gl, np.ascontiguousarray(cand, dtype=np.int32), nc


class _Tables:
    """Index tables of compute_all_dets' enumeration (aats.py:581-618), built by the C library
    (bit-exact contract) and uploaded once per (ndocc, nfzc, nvirt)."""

    def __init__(self, no, nf, nv):
        self.no, self.nf, self.nv = no, nf, nv
        ns, nd = C.c_int64(), C.c_int64()
        check(lib.apyib_det_enumeration(no, nf, nv, None, C.byref(ns), None, C.byref(nd)))
        self.n1, self.n2 = ns.value, nd.value
        self.singles = np.zeros((self.n1, 2), dtype=np.int32)
        self.doubles = np.zeros((self.n2, 4), dtype=np.int32)
        check(lib.apyib_det_enumeration(no, nf, nv, _i32_host(self.singles)[1], None,
                                        _i32_host(self.doubles)[1], None))
        self.L = []
        for sub, cnt, nsub in ((None, 1, 0), (self.singles, self.n1, 1), (self.doubles, self.n2, 2)):
            out = np.zeros((max(cnt, 1), no), dtype=np.int32)
            if cnt:
                check(lib.apyib_det_index_lists(no, _i32_host(sub)[1] if sub is not None else None, cnt, nsub,
                                                _i32_host(out)[1]))
            self.L.append(torch.from_numpy(out[:cnt].copy()).to(device()))
        # column-side lists re-ordered for factorisation reuse in the thread-per-matrix LU kernel
        # (substituted columns last, lists sorted; include/apyib_b200.h: apyib_det_sort_lists)
        self.LS = [None, None, None]
        self.PFX = [None, None, None]       # (group_len, candidate columns) when the sorted lists have the group structure
        if 2 <= no <= 12:
            for k, cnt in ((1, self.n1), (2, self.n2)):
                if not cnt:
                    continue
                src = self.L[k].cpu().numpy()
                srt = np.zeros_like(src)
                sign = np.zeros(cnt, dtype=np.float64)
                idx = np.zeros(cnt, dtype=np.int32)
                check(lib.apyib_det_sort_lists(no, _i32_host(src)[1], cnt, _i32_host(srt)[1],
                                               sign.ctypes.data_as(C.POINTER(C.c_double)), _i32_host(idx)[1]))
                self.LS[k] = tuple(torch.from_numpy(x.copy()).to(device()) for x in (srt, sign, idx))
                grp = prefix_groups(srt, no, k)
                if grp is not None:
                    self.PFX[k] = (grp[0], torch.from_numpy(grp[1]).to(device()), grp[2])
        self.doubles_dev = torch.from_numpy(self.doubles.copy()).to(device())
        self.singles_dev = torch.from_numpy(self.singles.copy()).to(device())

    @staticmethod
    def get(no, nf, nv):
        key = (no, nf, nv, torch.cuda.current_device())
        if key not in _tables:
            _tables[key] = _Tables(no, nf, nv)
        return _tables[key]


def _det_outer(S, n, rows, cols, sorted_cols=None):
    out = empty((rows.shape[0], cols.shape[0]), _C128)
    if sorted_cols is not None and config.LU_REUSE:
        cs, sg, ix = sorted_cols
        check(lib.apyib_det_outer_sorted(ptr(S), S.shape[0], n, ptr(rows), rows.shape[0], ptr(cs), ptr(sg), ptr(ix),
                                         cs.shape[0], ptr(out), stream_ptr()))
        return out
    check(lib.apyib_det_outer(ptr(S), S.shape[0], n, ptr(rows), rows.shape[0], ptr(cols), cols.shape[0], ptr(out),
                              stream_ptr()))
    return out


def _det_matvec(S, n, rows, cols, Y, sorted_cols=None, prefix=None, k=0):
    """Z[q, r] = sum_c det(S[rows[r], cols[c]]) Y[q, c]; any number of vectors (<= 4 per launch).
    prefix = (group_len, candidate columns, nc) of a k-fold substituted table routes the launch to the
    prefix-shared LU kernel (config.LU_PREFIX)."""
    use_sorted = sorted_cols is not None and config.LU_REUSE
    use_prefix = use_sorted and prefix is not None and config.LU_PREFIX
    Y = Y.contiguous()
    nq, nrow, ncol = Y.shape[0], rows.shape[0], cols.shape[0]
    Z = empty((nq, nrow), _C128)
    for q0 in range(0, nq, 4):
        q1 = min(nq, q0 + 4)
        if use_prefix:
            gl, cand, nc = prefix
            cs, sg, ix = sorted_cols
            work = empty((int(lib.apyib_det_matvec_pairs_work_len(nrow, ncol // gl, q1 - q0, n, k, S.shape[0], nc)),), _C128)
            with config.timed("det_pairs[n=%d,k=%d,%dx%d]" % (n, k, nrow, ncol)):
                rc = lib.apyib_det_matvec_pairs(ptr(S), S.shape[0], n, k, ptr(rows), nrow, ptr(cs), ptr(sg), ptr(ix), ncol, gl,
                                                ptr(cand), nc, ptr(Y[q0:q1]), q1 - q0, ptr(Z[q0:q1]), ptr(work), stream_ptr())
            if rc == 0:
                continue
            if rc != -3:          # APYIB_ERR_UNSUPPORTED -> the general kernel below
                check(rc)
        work = empty((int(lib.apyib_det_matvec_work_len(nrow, ncol, q1 - q0, n)),), _C128)
        with config.timed("det_matvec[n=%d,%dx%d]" % (n, nrow, ncol)):
            if use_sorted:
                cs, sg, ix = sorted_cols
                check(lib.apyib_det_matvec_sorted(ptr(S), S.shape[0], n, ptr(rows), nrow, ptr(cs), ptr(sg), ptr(ix), ncol,
                                                  ptr(Y[q0:q1]), q1 - q0, ptr(Z[q0:q1]), ptr(work), stream_ptr()))
            else:
                check(lib.apyib_det_matvec(ptr(S), S.shape[0], n, ptr(rows), nrow, ptr(cols), ncol, ptr(Y[q0:q1]),
                                           q1 - q0, ptr(Z[q0:q1]), ptr(work), stream_ptr()))
    return Z


def _dev(x):
    """numpy array / python scalar / CUDA tensor -> complex128 CUDA tensor"""
    if isinstance(x, torch.Tensor):
        return to_device(x, _C128)
    return to_device(np.asarray(x), _C128)


def _axpby(alpha, x, beta, y, conj_x=False):
    a, b = complex(alpha), complex(beta)
    check(lib.apyib_axpby(1, x.numel(), a.real, a.imag, ptr(x), int(conj_x), b.real, b.imag, ptr(y), stream_ptr()))
    return y


def _vdot(x, y, conj_x=True):
    """sum op(x) * y over equally laid out device tensors -> python complex"""
    out = zeros((2,), torch.float64)
    check(lib.apyib_dots(1, ptr(x), 0, 1, ptr(y), x.numel(), int(conj_x), ptr(out), ptr(reduce_scratch()), stream_ptr()))
    h = to_host(out)
    return complex(h[0], h[1])



# -------------------------------------------------------------------------------------------------
# CUDA-graph replay of the device part of AAT._blocks (config.AAT_USE_GRAPH)
# -------------------------------------------------------------------------------------------------
_block_graphs = {}          # stack shape -> _BlockGraph | "warm" (seen once, eager) | None (not capturable)
GRAPH_MAX_BYTES = 6 << 30
GRAPH_MAX_LIVE = 24
import os as _os
BLK4_GROUP = int(_os.environ.get("APYIB_B200_BLK4_GROUP", "5"))   # nuclear coordinates per (beta x pp/pn/np/nn) overlap stack: 12 overlaps each


class _BlockGraph:
    """`AAT._blocks_device` for one stack shape, captured over static input buffers.  Capture runs under
    torch.cuda.graph so that the intermediates the launch sequence allocates come from a private pool that
    stays reserved for the graph (their addresses are baked into the captured launches)."""

    def __init__(self, aat, inputs):
        self.inp = [None if t is None else t.clone() for t in inputs]
        self.graph = torch.cuda.CUDAGraph()
        n0 = _liblaunch[0]
        # thread_local: other threads (NCCL watchdog, bench clock sampler) must not invalidate the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.flat, self.meta = aat._blocks_device(*self.inp)
        self.launches = _liblaunch[0] - n0
        _liblaunch[0] = n0                                   # recorded, not executed

    def run(self, inputs):
        for dst, src in zip(self.inp, inputs):
            if dst is not None:
                dst.copy_(src)
        self.graph.replay()
        _liblaunch[0] += self.launches
        return self.flat, self.meta


def _run_blocks(aat, S, X1, X2, Y1, Y2):
    inputs = (S, X1, X2, Y1, Y2)
    if not config.AAT_USE_GRAPH or torch.cuda.is_current_stream_capturing():
        return aat._blocks_device(*inputs)
    o, nv = aat.ndocc - aat.nfzc, aat.nbf - aat.ndocc
    nS, nx, ny = S.shape[0], X2.shape[0], Y2.shape[0]
    if 16 * o * o * nv * nv * nS * (nx + ny + nx * ny) > GRAPH_MAX_BYTES:
        return aat._blocks_device(*inputs)                   # large shapes are not launch-bound; keep the pool small
    key = (config.AAT_ALGORITHM, config.LU_REUSE, config.LU_PREFIX, config.USE_TMA, aat.nbf, aat.ndocc, aat.nfzc, nS, nx, ny, X1 is not None,
           tuple(X2.shape), tuple(Y2.shape), torch.cuda.current_device())
    g = _block_graphs.get(key, "new")
    if g == "new":
        # first sight of a shape: eager (builds the contraction offset tables and one-time function attributes)
        _block_graphs[key] = "warm"
        return aat._blocks_device(*inputs)
    if g == "warm":
        live = sum(1 for v in _block_graphs.values() if isinstance(v, _BlockGraph))
        if live >= GRAPH_MAX_LIVE:                           # bound the memory held by private graph pools:
            _block_graphs[key] = None                        # further shapes simply stay eager
            return aat._blocks_device(*inputs)
        try:
            g = _BlockGraph(aat, inputs)
        except Exception as exc:                             # not capturable here: stay eager for this shape
            if config.VERBOSE:
                print("apyib_b200: CUDA-graph capture of the AAT block failed (%s); running eagerly" % exc)
            g = None
        _block_graphs[key] = g
    if g is None:
        return aat._blocks_device(*inputs)
    return g.run(inputs)


class AAT(object):
    """The atomic axial tensor object computed by finite difference (aats.py:19-115)."""

    def __init__(self, parameters, wfn, unperturbed_wfn, unperturbed_basis, unperturbed_T, nuc_pos_wfn, nuc_neg_wfn,
                 nuc_pos_basis, nuc_neg_basis, nuc_pos_T, nuc_neg_T, mag_pos_wfn, mag_neg_wfn, mag_pos_basis,
                 mag_neg_basis, mag_pos_T, mag_neg_T, nuc_pert_strength, mag_pert_strength, rows=None):
        """`rows` (optional, not in the reference): the nuclear coordinates alpha this process is going to evaluate
        (sharded driver, parallel.py).  Overlaps, norms and scaled amplitudes of other rows are then never built;
        their list entries (C / basis / T) may be None."""
        from .hostchem import provider_ao_overlap
        self.nuc_pos_wfn, self.nuc_neg_wfn = nuc_pos_wfn, nuc_neg_wfn
        self.nuc_pos_T, self.nuc_neg_T = nuc_pos_T, nuc_neg_T
        self.mag_pos_wfn, self.mag_neg_wfn = mag_pos_wfn, mag_neg_wfn
        self.mag_pos_T, self.mag_neg_T = mag_pos_T, mag_neg_T
        self.nuc_pert_strength, self.mag_pert_strength = nuc_pert_strength, mag_pert_strength
        self.unperturbed_wfn, self.unperturbed_T = unperturbed_wfn, unperturbed_T
        natom = len(nuc_pos_wfn) // 3
        self.nbf, self.ndocc = wfn.nbf, wfn.ndocc
        self.nfzc = wfn.H.basis_set.n_frozen_core()
        self.parameters = parameters
        so = parameters["method"] in SO_METHODS

        # utils.compute_mo_overlap (+ compute_so_overlap), utils.py:370-422, for all 1 + 6 + 6N + 36N (bra, ket)
        # pairs at once: the AO overlaps are host inputs, the C^H S C products two batched launches per dtype
        jobs = []

        ao_cache, geo_cache = {}, {}

        def geo_key(basis):
            k = geo_cache.get(id(basis))
            if k is None:
                try:
                    k = (id(basis.provider), basis.molecule.geometry().tobytes())
                except AttributeError:
                    k = id(basis)
                geo_cache[id(basis)] = k
            return k

        def ao_overlap(bb, kb):
            """mixed-geometry AO overlap, evaluated once per distinct (bra geometry, ket geometry): all field points
            share the unperturbed geometry, so only 6N + 1 of the 42N + 7 calls are different"""
            key = (geo_key(bb), geo_key(kb))
            if key not in ao_cache:
                ao_cache[key] = provider_ao_overlap(bb, kb)
            return ao_cache[key]

        def ovl(bb, Cb, kb, Ck):
            jobs.append((Cb, ao_overlap(bb, kb), Ck))
            return len(jobs) - 1

        U, Ub = self.unperturbed_wfn, unperturbed_basis
        n3 = 3 * natom
        rhf = parameters["method"] == "RHF"
        if rows is not None:
            self._rows = sorted(set(int(a) for a in rows))
        mine = set(range(n3)) if rows is None else set(self._rows)
        _ovl = ovl
        ovl = lambda bb, Cb, kb, Ck: _ovl(bb, Cb, kb, Ck) if Cb is not None else None      # row not owned
        sel = lambda lst, a: lst[a] if a in mine else None
        if not rhf:
            uu = ovl(Ub, U, Ub, U)
            up = [ovl(Ub, U, mag_pos_basis[b], mag_pos_wfn[b]) for b in range(3)]
            un = [ovl(Ub, U, mag_neg_basis[b], mag_neg_wfn[b]) for b in range(3)]
            pu = [ovl(nuc_pos_basis[a], sel(nuc_pos_wfn, a), Ub, U) for a in range(n3)]
            nu = [ovl(nuc_neg_basis[a], sel(nuc_neg_wfn, a), Ub, U) for a in range(n3)]
        pp = [[ovl(nuc_pos_basis[a], sel(nuc_pos_wfn, a), mag_pos_basis[b], mag_pos_wfn[b]) for b in range(3)] for a in range(n3)]
        pn = [[ovl(nuc_pos_basis[a], sel(nuc_pos_wfn, a), mag_neg_basis[b], mag_neg_wfn[b]) for b in range(3)] for a in range(n3)]
        np_ = [[ovl(nuc_neg_basis[a], sel(nuc_neg_wfn, a), mag_pos_basis[b], mag_pos_wfn[b]) for b in range(3)] for a in range(n3)]
        nn = [[ovl(nuc_neg_basis[a], sel(nuc_neg_wfn, a), mag_neg_basis[b], mag_neg_wfn[b]) for b in range(3)] for a in range(n3)]
        if so:
            host = [to_host(spin_block_2_dev(S)) for S in mo_overlaps_dev(jobs)]
        else:
            host = mo_overlaps_dev(jobs, host=True)
        get = lambda x: [get(y) for y in x] if isinstance(x, list) else (None if x is None else host[x])
        if not rhf:
            self.overlap_uu, self.overlap_up, self.overlap_un = get(uu), get(up), get(un)
            self.overlap_pu, self.overlap_nu = get(pu), get(nu)
        self.overlap_pp, self.overlap_pn, self.overlap_np, self.overlap_nn = get(pp), get(pn), get(np_), get(nn)
        self._cache = {}

    @classmethod
    def from_parts(cls, method, nbf, ndocc, nfzc, h_R, h_B, **parts):
        """Build from precomputed overlaps / amplitude lists (what __init__ leaves on `self`);
        used by tests and by the sharded driver, which computes overlaps rank-locally."""
        self = cls.__new__(cls)
        self.parameters = {"method": method}
        self.nbf, self.ndocc, self.nfzc = nbf, ndocc, nfzc
        self.nuc_pert_strength, self.mag_pert_strength = h_R, h_B
        for k, v in parts.items():
            setattr(self, k, v)
        self._cache = {}
        return self

    # ---------------------------------------------------------------------------------------
    # aats.py:120-130
    def compute_SO_det(self, overlap, bra_indices, ket_indices):
        nocc = 2 * self.ndocc
        S = to_device(np.asarray(overlap), _C128)
        nso = S.shape[0]
        lists = []
        for idx in (bra_indices, ket_indices):
            out = np.zeros((1, nocc), dtype=np.int32)
            sub, p = _i32_host(np.asarray(idx, dtype=np.int32).reshape(-1))
            check(lib.apyib_so_index_lists(nso, nocc, p if len(idx) else None, 1, len(idx) // 2, _i32_host(out)[1]))
            lists.append(torch.from_numpy(out).to(device()))
        return complex(to_host(_det_outer(S, nocc, lists[0], lists[1]))[0, 0])

    # ---------------------------------------------------------------------------------------
    # aats.py:134-157
    def compute_normalization(self, alpha, beta, normalization):
        m = self.parameters["method"]
        if m == "RHF" or normalization == "intermediate":
            return 1, 1, 1, 1, 1
        cisd = m == "CISD_SO"

        def N(T):
            t2 = _dev(T[2])
            x = T[0] ** 2 + 0.25 * _vdot(t2, t2)
            if cisd:
                t1 = _dev(T[1])
                x = x + _vdot(t1, t1)
            return 1 / np.sqrt(x)

        return (N(self.unperturbed_T), N(self.nuc_pos_T[alpha]), N(self.nuc_neg_T[alpha]),
                N(self.mag_pos_T[beta]), N(self.mag_neg_T[beta]))

    # ---------------------------------------------------------------------------------------
    # aats.py:558-642
    def compute_all_dets(self, overlap):
        """All singly / doubly row- and/or column-substituted determinants of one MO overlap, in the
        reference's nine return objects (materialised; use compute_spatial_aats for the fused path)."""
        no, nf, nv = self.ndocc, self.nfzc, self.nbf - self.ndocc
        o = no - nf
        T = _Tables.get(no, nf, nv)
        S = to_device(np.asarray(overlap), _C128)
        R0, R1, R2 = T.L
        d = lambda r, c: to_host(_det_outer(S, no, r, c))
        det_S = complex(d(R0, R0)[0, 0])
        si, sa = T.singles[:, 0] - nf, T.singles[:, 1]
        di, da, dj, db = T.doubles[:, 0] - nf, T.doubles[:, 1], T.doubles[:, 2] - nf, T.doubles[:, 3]
        z = lambda *s: np.zeros(s, dtype=np.complex128)
        ia_S, S_kc = z(o, nv), z(o, nv)
        iajb_S, S_kcld, ia_S_kc = z(o, nv, o, nv), z(o, nv, o, nv), z(o, nv, o, nv)
        iajb_S_kc, ia_S_kcld = z(o, nv, o, nv, o, nv), z(o, nv, o, nv, o, nv)
        iajb_S_kcld = z(*((o, nv) * 4))
        if T.n1:
            ia_S[si, sa] = d(R1, R0)[:, 0]
            S_kc[si, sa] = d(R0, R1)[0, :]
            ia_S_kc[si[:, None], sa[:, None], si[None, :], sa[None, :]] = d(R1, R1)
        if T.n2:
            iajb_S[di, da, dj, db] = d(R2, R0)[:, 0]
            S_kcld[di, da, dj, db] = d(R0, R2)[0, :]
            iajb_S_kc[di[:, None], da[:, None], dj[:, None], db[:, None], si[None, :], sa[None, :]] = d(R2, R1)
            # stored [i][a][j][b][k][c] with COLUMNS (i,a),(j,b) and ROW (k,c) substituted (aats.py:604-606)
            ia_S_kcld[di[None, :], da[None, :], dj[None, :], db[None, :], si[:, None], sa[:, None]] = d(R1, R2)
            iajb_S_kcld[di[:, None], da[:, None], dj[:, None], db[:, None],
                        di[None, :], da[None, :], dj[None, :], db[None, :]] = d(R2, R2)
        a4 = lambda X: X - X.swapaxes(0, 2) - X.swapaxes(1, 3) + X.swapaxes(0, 2).swapaxes(1, 3)
        iajb_S, S_kcld, iajb_S_kc, ia_S_kcld = a4(iajb_S), a4(S_kcld), a4(iajb_S_kc), a4(ia_S_kcld)
        ia_S_kcld = ia_S_kcld.swapaxes(0, 4).swapaxes(1, 5).swapaxes(2, 4).swapaxes(3, 5)       # aats.py:629
        X = a4(iajb_S_kcld)
        iajb_S_kcld = X - X.swapaxes(4, 6) - X.swapaxes(5, 7) + X.swapaxes(4, 6).swapaxes(5, 7)  # aats.py:630
        return det_S, ia_S, S_kc, iajb_S, S_kcld, ia_S_kc, iajb_S_kc, ia_S_kcld, iajb_S_kcld

    # ---------------------------------------------------------------------------------------
    # spatial route
    # ---------------------------------------------------------------------------------------
    def _amp(self, x):
        """complex128 device copy of an amplitude array, uploaded / converted once per AAT object (the norms and
        the scaled amplitude combinations both read every amplitude set)"""
        memo = self.__dict__.setdefault("_amp_memo", {})
        ent = memo.get(id(x))
        if ent is None or ent[0] is not x:
            ent = memo[id(x)] = (x, _dev(x))
        return ent[1]

    def _active_rows(self):
        """nuclear coordinates whose norms / scaled amplitudes are built: the hinted rows (constructor `rows=`,
        prefetch_rows) or all of them"""
        n3 = len(self.nuc_pos_T)
        rows = getattr(self, "_rows", None)
        return list(range(n3)) if not rows else list(rows)

    def _spatial_norms(self, normalization):
        """N, N_np[3N], N_nn[3N], N_mp[3], N_mn[3] of aats.py:652-669 (entries of rows that are not active: None)."""
        rows = self._active_rows()
        key = ("norms", normalization, tuple(rows))
        if key in self._cache:
            return self._cache[key]
        m = self.parameters["method"]
        n3 = len(self.nuc_pos_T)
        if m == "RHF" or normalization == "intermediate":
            res = (1, [1] * n3, [1] * n3, [1] * 3, [1] * 3)
        else:
            cisd = m == "CISD"
            # all needed points as one stack: three batched dot-like contractions and ONE device->host copy
            # (x = t0 + 2<t1|t1> + 2<t2|t2> - <t2|t2^T>, aats.py:652-669)
            Ts = ([self.unperturbed_T] + [self.nuc_pos_T[a] for a in rows] + [self.nuc_neg_T[a] for a in rows]
                  + list(self.mag_pos_T) + list(self.mag_neg_T))
            t2 = torch.stack([self._amp(T[2]) for T in Ts])
            npt = len(Ts)
            acc = zeros((3, npt), _C128)
            contract("sijab,sijab->s", t2, t2, acc[0], 1.0, 0.0, conj_a=True)
            contract("sijab,sijba->s", t2, t2, acc[1], 1.0, 0.0, conj_a=True)
            if cisd:
                t1 = torch.stack([self._amp(T[1]) for T in Ts])
                contract("sia,sia->s", t1, t1, acc[2], 1.0, 0.0, conj_a=True)
            h = to_host(acc)
            Nall = [1 / np.sqrt(Ts[s][0] + (2 * complex(h[0, s]) - complex(h[1, s])) + (2 * complex(h[2, s]) if cisd else 0))
                    for s in range(npt)]
            nr = len(rows)
            N_np, N_nn = [None] * n3, [None] * n3
            for k, a in enumerate(rows):
                N_np[a], N_nn[a] = Nall[1 + k], Nall[1 + nr + k]
            res = (Nall[0], N_np, N_nn, Nall[1 + 2 * nr:4 + 2 * nr], Nall[4 + 2 * nr:7 + 2 * nr])
        self._cache[key] = res
        return res

    def _spatial_amps(self, normalization):
        """Scaled amplitude combinations of aats.py:690-711 for the active alpha / all beta, on the device:
        kets  Y_t (1), Y_dH (3);  bras X_c (1), X_dR (active rows; "pos" maps alpha -> position in the stack)."""
        rows = self._active_rows()
        key = ("amps", normalization, tuple(rows))
        if key in self._cache:
            return self._cache[key]
        cisd = self.parameters["method"] == "CISD"
        N, N_np, N_nn, N_mp, N_mn = self._spatial_norms(normalization)

        def build(idx):
            if idx == 1 and not cisd:
                return None
            T0 = self._amp(self.unperturbed_T[idx])
            t = _axpby(N, T0, 0.0, torch.empty_like(T0))
            tc = _axpby(np.conj(N), T0, 0.0, torch.empty_like(T0), conj_x=True)
            dH, dR = [], []
            for b in range(3):
                x = _axpby(N_mp[b], self._amp(self.mag_pos_T[b][idx]), 0.0, torch.empty_like(T0))
                dH.append(_axpby(-N_mn[b], self._amp(self.mag_neg_T[b][idx]), 1.0, x))
            for a in rows:
                x = _axpby(np.conj(N_np[a]), self._amp(self.nuc_pos_T[a][idx]), 0.0,
                           torch.empty_like(T0), conj_x=True)
                dR.append(_axpby(-np.conj(N_nn[a]), self._amp(self.nuc_neg_T[a][idx]), 1.0, x,
                                 conj_x=True))
            return dict(t=t[None], tc=tc[None], dH=torch.stack(dH), dR=torch.stack(dR))

        res = {1: build(1), 2: build(2), "pos": {a: k for k, a in enumerate(rows)}}
        self._cache[key] = res
        self.__dict__.pop("_amp_memo", None)           # the raw device copies are not needed any more
        return res

    def _block(self, S_host, X1, X2, Y1, Y2):
        return self._blocks([S_host], X1, X2, Y1, Y2)[0]

    def _blocks(self, S_hosts, X1, X2, Y1, Y2):
        """Contributions of a STACK of overlap matrices for nx bra and ny ket amplitude sets, all
        launches batched over the stack.  Returns one dict per overlap of [nx, ny] numpy arrays
        (before the +/- sign and the N factors of S0/0S).  config.AAT_ALGORITHM selects how the
        substituted determinants are evaluated: "lu" (LU of every n x n matrix, csrc/dets_tpm.cu /
        dets.cu), "lemma" (<= 4 x 4 determinants from S_oo^-1, csrc/lemma.cu) or "factorized".

        The device part (`_blocks_device`: ~15 launches per overlap on the LU path) is a fixed launch
        sequence for a given stack shape, so with config.AAT_USE_GRAPH it is captured once per shape as a
        CUDA graph over static input buffers and replayed: the 12 (beta x pp/pn/np/nn) stacks of every
        nuclear coordinate, and every later molecule of the same shape, cost one graph launch instead of
        ~200 host-issued launches (the small-molecule assembly is launch-bound otherwise)."""
        S = to_device(np.stack([np.asarray(x) for x in S_hosts]).astype(np.complex128), _C128)     # [nS, ns, ns]
        nS, nx, ny = len(S_hosts), X2.shape[0], Y2.shape[0]
        cisd = X1 is not None
        flat, meta = _run_blocks(self, S, X1, X2, Y1, Y2)
        return self._blocks_host(to_host(flat), meta, nS, nx, ny, cisd)

    def _blocks_device(self, S, X1, X2, Y1, Y2):
        """Device part of `_blocks`: returns (flat, meta) -- one flat complex128 device tensor holding det_S
        of every overlap followed by all small result tensors, and meta = [(key, shape), ...]."""
        no, nf, nv = self.ndocc, self.nfzc, self.nbf - self.ndocc
        o = no - nf
        cisd = X1 is not None
        lemma = config.AAT_ALGORITHM in ("lemma", "factorized")
        factorized = config.AAT_ALGORITHM == "factorized"
        T = _Tables.get(no, nf, nv)
        R0, R1, R2 = T.L
        nS, ns = S.shape[0], self.nbf
        nx, ny = X2.shape[0], Y2.shape[0]
        P, n1 = T.n2, T.n1
        cn = contract_new
        X2, Y2 = X2.contiguous(), Y2.contiguous()
        if lemma:
            prep = empty((nS, int(lib.apyib_lemma_prep_len(ns, no))), _C128)
            check(lib.apyib_lemma_prepare(ptr(S), nS, ns, no, ptr(prep), stream_ptr()))
            subs = {0: None, 1: T.singles_dev, 2: T.doubles_dev}
            cnt = {0: 1, 1: n1, 2: P}

            def outer(rk, ck):
                out = empty((nS, cnt[rk], cnt[ck]), _C128)
                check(lib.apyib_lemma_outer(ptr(prep), nS, ns, no, rk, ptr(subs[rk]), cnt[rk], ck, ptr(subs[ck]),
                                            cnt[ck], ptr(out), stream_ptr()))
                return out

            def matvec(rk, ck, Y, per_overlap):
                """Z[s,q,r] = sum_c D_s[r,c] Y[(s,)q,c]"""
                Y = Y.contiguous()
                nq = Y.shape[-2]
                Z = empty((nS, nq, cnt[rk]), _C128)
                for q0 in range(0, nq, 4):
                    q1 = min(nq, q0 + 4)
                    Yq = Y[..., q0:q1, :].contiguous()
                    Zq = Z if (q0 == 0 and q1 == nq) else empty((nS, q1 - q0, cnt[rk]), _C128)
                    work = empty((int(lib.apyib_lemma_matvec_work_len(cnt[rk], cnt[ck], q1 - q0, nS)),), _C128)
                    with config.timed("lemma_matvec[%d%d,%dx%d,nS=%d]" % (rk, ck, cnt[rk], cnt[ck], nS)):
                        check(lib.apyib_lemma_matvec(ptr(prep), nS, ns, no, rk, ptr(subs[rk]), cnt[rk], ck,
                                                     ptr(subs[ck]), cnt[ck], ptr(Yq),
                                                     (q1 - q0) * cnt[ck] if per_overlap else 0, q1 - q0, ptr(Zq),
                                                     ptr(work), stream_ptr()))
                    if Zq is not Z:
                        Z[:, q0:q1].copy_(Zq)
                return Z
        else:
            L = {0: R0, 1: R1, 2: R2}
            cnt = {0: 1, 1: n1, 2: P}
            null = C.c_void_p(0)

            def sorted_lists(ck):
                """(column lists, sign, index) for table kind ck: re-ordered for factorisation reuse when available"""
                if T.LS[ck] is not None and config.LU_REUSE:
                    cs, sg, ix = T.LS[ck]
                    return ptr(cs), ptr(sg), ptr(ix), True
                return ptr(L[ck]), null, null, False

            def outer(rk, ck):
                """D[s, r, c] for the whole stack in one launch (grid.y = overlap)"""
                out = empty((nS, cnt[rk], cnt[ck]), _C128)
                if cnt[rk] and cnt[ck]:
                    cols, sg, ix, _ = sorted_lists(ck)
                    check(lib.apyib_det_outer_stack(ptr(S), nS, ns, no, ptr(L[rk]), cnt[rk], cols, sg, ix, cnt[ck], ptr(out),
                                                    stream_ptr()))
                return out

            def matvec(rk, ck, Y, per_overlap):
                """Z[s,q,r] = sum_c D_s[r,c] Y[(s,)q,c] for the whole stack, <= 4 vectors per launch"""
                Y = Y.contiguous()
                nq, nrow, ncol = Y.shape[-2], cnt[rk], cnt[ck]
                Z = empty((nS, nq, nrow), _C128)
                cols, sg, ix, is_sorted = sorted_lists(ck)
                pfx = T.PFX[ck] if (is_sorted and config.LU_PREFIX) else None
                for q0 in range(0, nq, 4):
                    q1 = min(nq, q0 + 4)
                    whole = q0 == 0 and q1 == nq
                    Yq = Y if whole else Y[..., q0:q1, :].contiguous()
                    Zq = Z if whole else empty((nS, q1 - q0, nrow), _C128)
                    ystride = (q1 - q0) * ncol if per_overlap else 0
                    rc = -3
                    if pfx is not None:
                        gl, cand, nc = pfx
                        work = empty((nS * int(lib.apyib_det_matvec_pairs_work_len(nrow, ncol // gl, q1 - q0, no, ck, ns, nc)),), _C128)
                        with config.timed("det_pairs[n=%d,k=%d,%dx%d,nS=%d]" % (no, ck, nrow, ncol, nS)):
                            rc = lib.apyib_det_matvec_pairs_stack(ptr(S), nS, ns, no, ck, ptr(L[rk]), nrow, cols, sg, ix, ncol, gl,
                                                                  ptr(cand), nc, ptr(Yq), ystride, q1 - q0, ptr(Zq), ptr(work),
                                                                  stream_ptr())
                        if rc not in (0, -3):      # -3 = APYIB_ERR_UNSUPPORTED -> the per-matrix LU below
                            check(rc)
                    if rc != 0:
                        work = empty((nS * int(lib.apyib_det_matvec_work_len(nrow, ncol, q1 - q0, no)),), _C128)
                        with config.timed("det_matvec[n=%d,%dx%d,nS=%d]" % (no, nrow, ncol, nS)):
                            check(lib.apyib_det_matvec_stack(ptr(S), nS, ns, no, ptr(L[rk]), nrow, cols, sg, ix, ncol, ptr(Yq),
                                                             ystride, q1 - q0, ptr(Zq), ptr(work), stream_ptr()))
                    if Zq is not Z:
                        Z[:, q0:q1].copy_(Zq)
                return Z

        dS = outer(0, 0).reshape(nS)                                     # det_S
        A = outer(1, 0).reshape(nS, o, nv)                               # ia_S
        B = outer(0, 1).reshape(nS, o, nv)                               # S_kc
        G = outer(1, 1).reshape(nS, o, nv, o, nv)                        # ia_S_kc
        dd = {}
        wy = cn("qklcd,sld->sqkc", Y2, B)                                # sum_ld y2[k,l,c,d] S_ld
        ux = cn("xijab,sjb->sxia", X2, A)                                # sum_jb x2[i,j,a,b] jb_S
        uxp = cn("xijab,sia->sxjb", X2, A)                               # sum_ia x2[i,j,a,b] ia_S
        if not (factorized and P):
            T1 = cn("siakc,qklcd->sqiald", G, Y2)
            Z = cn("sqiald,sjbld->sqiajb", T1, G)
            dd["c6"] = cn("xijab,sqiajb->sxq", X2, Z)                    # sum x2 y2 ia_S_kc jb_S_ld   (aats.py:737)
        if P:
            D20 = outer(2, 0).reshape(nS, P)                             # iajb_S  (restricted)
            D02 = outer(0, 2).reshape(nS, P)                             # S_kcld  (restricted)
            Xh, Yh = empty((nx, P), _C128), empty((ny, P), _C128)
            n2 = o * o * nv * nv
            check(lib.apyib_pack_doubles(ptr(X2), n2, nx, o, nv, nf, ptr(T.doubles_dev), P, ptr(Xh), stream_ptr()))
            check(lib.apyib_pack_doubles(ptr(Y2), n2, ny, o, nv, nf, ptr(T.doubles_dev), P, ptr(Yh), stream_ptr()))
            uxs = _axpby(1.0, uxp, 1.0, ux.clone())                      # ux + ux'
            dd["v1"] = cn("xr,sr->sx", Xh, D20)
            dd["v2"] = cn("sr,qr->sq", D02, Yh)
            if factorized:
                # doubles x doubles, doubles x singles and singles x doubles tables in closed form
                f = self._dd_factorized(prep, nS, ns, no, nf, X2, Y2)
                dd["c1f"] = f["dd"]
                dd["c6f"] = f["c6"]
                half = lambda a, b, rest: _axpby(0.5, cn("sx,sq->sxq", a, b), 1.0, rest)
                # sum_r Xh[r] D21[r,c] y[c] / det(S_oo) = alpha (Q.y)/2 + M.y
                dd["c3f"] = half(f["alpha"], cn("skc,sqkc->sq", f["Q"], wy), cn("sxkc,sqkc->sxq", f["M"], wy))
                # sum_c D12[(i,a),c] Yh[c] / det(S_oo) = beta P_ai/2 + N_ia,  N = A^-T W R^T
                Nq = cn("ski,sqka->sqia", f["Ai"], cn("sac,sqkc->sqka", f["R"], f["W"]))
                dd["c4f"] = half(cn("sxia,sai->sx", uxs, f["P"]), f["beta"], cn("sxia,sqia->sxq", uxs, Nq))
                if cisd:
                    dd["ds1f"] = half(f["alpha"], cn("skc,qkc->sq", f["Q"], Y1), cn("sxkc,qkc->sxq", f["M"], Y1))
                    dd["sd1f"] = half(cn("xia,sai->sx", X1, f["P"]), f["beta"], cn("xia,sqia->sxq", X1, Nq))
            else:
                z22 = matvec(2, 2, Yh, False)                            # the P x P table, fused      [s,q,r]
                ys = [wy.reshape(nS, ny, -1)] + ([Y1.reshape(1, ny, -1).expand(nS, ny, n1)] if cisd else [])
                z21 = matvec(2, 1, torch.cat(ys, 1).contiguous(), True)  # [s, ny(+ny), P]
                z12 = matvec(1, 2, Yh, False)                            # [s, q, ov]
                dd["c1"] = cn("xr,sqr->sxq", Xh, z22)
                dd["c3"] = cn("xr,sqr->sxq", Xh, z21[:, :ny])
                dd["c4"] = cn("sxr,sqr->sxq", uxs.reshape(nS, nx, -1), z12)
                if cisd:
                    dd["ds1"] = cn("xr,sqr->sxq", Xh, z21[:, ny:])
                    dd["sd1"] = cn("xr,sqr->sxq", X1.reshape(nx, -1), z12)
        if cisd:
            Gy1 = cn("siakc,qkc->sqia", G, Y1)
            Gwy = cn("siakc,sqkc->sqia", G, wy)
            dd["s_xGy"] = cn("xia,sqia->sxq", X1, Gy1)
            dd["s_xA"] = cn("xia,sia->sx", X1, A)
            dd["s_yB"] = cn("skc,qkc->sq", B, Y1)
            dd["ds3"] = cn("sxia,sqia->sxq", ux, Gy1)
            dd["sd3"] = cn("xia,sqia->sxq", X1, Gwy)
            dd["d0b"] = cn("sxjb,sjb->sx", uxp, A)
            dd["0db"] = cn("skc,sqkc->sq", B, wy)
        # one flat tensor -> one device->host copy for all the small result tensors of the stack
        keys = list(dd)
        flat = torch.cat([dS.reshape(-1)] + [dd[k].reshape(-1) for k in keys])
        return flat, [(k, tuple(dd[k].shape)) for k in keys]

    @staticmethod
    def _blocks_host(fh, meta, nS, nx, ny, cisd):
        """Host part of `_blocks`: signs, det(S_oo) factors and the assembly of the per-overlap terms."""
        dSh_all = fh[:nS]
        h, off = {}, nS
        for k, shape in meta:
            n_el = int(np.prod(shape)) if len(shape) else 1
            h[k] = fh[off:off + n_el].reshape(shape)
            off += n_el
        for k in ("c3", "c4", "ds1", "sd1"):          # closed forms are per det(S_oo), like c1f
            if k + "f" in h:
                h[k] = dSh_all.reshape(nS, 1, 1) * h[k + "f"]
        if "c6f" in h:                                # product of two singly-singly substituted determinants
            h["c6"] = (dSh_all ** 2).reshape(nS, 1, 1) * h["c6f"]
        res = []
        for s in range(nS):
            dSh = complex(dSh_all[s])
            zero = np.zeros((nx, ny), dtype=np.complex128)
            g = lambda k: h[k][s] if k in h else zero
            v1 = g("v1").reshape(nx, 1) if "v1" in h else np.zeros((nx, 1), dtype=np.complex128)
            v2 = g("v2").reshape(1, ny) if "v2" in h else np.zeros((1, ny), dtype=np.complex128)
            c1 = g("c1")
            if "c1f" in h:                    # factorised: 16 x (restricted double sum) / det(S_oo)
                c1 = dSh * h["c1f"][s] / 16.0
            out = {"dS": dSh, "DD": 0.125 * (dSh * c1 + v1 * v2 + 4 * g("c3") + 2 * g("c4") + 8 * g("c6"))}
            if cisd:
                xA, yB = g("s_xA").reshape(nx, 1), g("s_yB").reshape(1, ny)
                out["SS"] = 2 * (dSh * g("s_xGy") + xA * yB)
                out["DS"] = 0.5 * dSh * g("ds1") + 0.5 * v1 * yB + 2 * g("ds3")
                out["SD"] = 0.5 * dSh * g("sd1") + 0.5 * xA * v2 + 2 * g("sd3")
                out["S0"] = 2 * xA * dSh + zero
                out["0S"] = 2 * yB * dSh + zero
                out["D0"] = 0.5 * v1 * dSh + g("d0b").reshape(nx, 1) + zero
                out["0D"] = 0.5 * v2 * dSh + g("0db").reshape(1, ny) + zero
            res.append(out)
        return res

    def _dd_factorized(self, prep, nS, ns, no, nf, X2, Y2):
        """sum_{i,j,a,b,k,l,c,d} Xf[ijab] Yf[klcd] det T / det-free form of the doubles-doubles table
        contraction (SURVEY.md 8(f).1 "factorise the DD sums").  With the 4 x 4 lemma matrix
            T = [[P_ai P_aj R_ac R_ad], [P_bi P_bj R_bc R_bd], [A_ki A_kj -Q_kc -Q_kd], [A_li A_lj -Q_lc -Q_ld]]
        (A = S_oo^-1) expanded by 2 x 2 minors, and Xf / Yf the fully antisymmetrised amplitudes
        (antisymmetric under i<->j, a<->b resp. k<->l, c<->d), the 24 terms of det T collapse to
            sum Xf Yf det T = 4 [ (Xf.PP)(Yf.QQ) + 4 sum_kc (A U R)_kc W_kc + sum_klcd Z_klcd Yf_klcd ]
            U_jb = sum_ia Xf_ijab P_ai,  W_kc = sum_ld Yf_klcd Q_ld,  Z_klcd = sum A_ki A_lj R_ac R_bd Xf_ijab
        i.e. O(o^2 v^2 (o+v)) DMMA contractions instead of (C(o,2) C(v,2))^2 determinants.
        Returns the intermediates and "dd" [nS, nx, ny]: the unrestricted sum (= 16 x the restricted
        i<j,a<b,k<l,c<d sum) over det T."""
        from .utils import gather4
        nv = ns - no
        o = no - nf
        cn = contract_new
        body = prep[:, 1:]
        off = 0
        Ai = body[:, off:off + no * no].reshape(nS, no, no); off += no * no
        P = body[:, off:off + nv * no].reshape(nS, nv, no); off += nv * no
        Q = body[:, off:off + no * nv].reshape(nS, no, nv); off += no * nv
        R = body[:, off:off + nv * nv].reshape(nS, nv, nv)
        Ai, P, Q = Ai[:, nf:, nf:], P[:, :, nf:], Q[:, nf:, :]            # frozen-core rows/cols are never substituted

        def antisym(X):      # Xf[x,i,j,a,b] = 2 (x_ijab - x_ijba - x_jiab + x_jiba)
            out = torch.empty_like(X)
            for q in range(X.shape[0]):
                t = gather4(X[q], 0, X[q].shape, [0, 1, 2, 3], [0, 0, 0, 0], 1.0, [0, 1, 3, 2], [0, 0, 0, 0], -1.0)
                gather4(t, 0, t.shape, [0, 1, 2, 3], [0, 0, 0, 0], 2.0, [1, 0, 2, 3], [0, 0, 0, 0], -2.0, out=out[q])
            return out

        Xf, Yf = antisym(X2), antisym(Y2)
        U = cn("xijab,sai->sxjb", Xf, P)
        alpha = cn("sxjb,sbj->sx", U, P)
        W = cn("qklcd,sld->sqkc", Yf, Q)
        beta = cn("sqkc,skc->sq", W, Q)
        M = cn("sxkb,sbc->sxkc", cn("skj,sxjb->sxkb", Ai, U), R)
        mixed = cn("sxkc,sqkc->sxq", M, W)
        # Z = A A R R x2 is linear in the amplitudes and commutes with the antisymmetric completion
        # (Z[Xf] = antisym(Z[x2])), so the four o^2 v^2 (o + v) transforms are done ONCE, on the plain amplitudes,
        # and serve both the doubles x doubles table (with Xf, Yf) and the singly x singly product term below.
        # The quadrilinear form sum x_ijab y_klcd A_ki A_lj R_ac R_bd can be transformed from either side: the side
        # with FEWER amplitude sets is transformed (up/un: 30 bra sets against 1 ket set per overlap).
        nx, ny = X2.shape[0], Y2.shape[0]
        if ny < nx:
            Zy = cn("sbd,qklcd->sqklcb", R, Y2)
            Zy = cn("sac,sqklcb->sqklab", R, Zy)
            Zy = cn("slj,sqklab->sqkjab", Ai, Zy)
            Zy = cn("ski,sqkjab->sqijab", Ai, Zy)
            gamma = cn("xijab,sqijab->sxq", Xf, Zy, alpha=8.0)        # sum antisym(Zy).Xf = 8 sum Zy.Xf (see below)
            zxy = cn("xijab,sqijab->sxq", X2, Zy)
        else:
            Zx = cn("ski,xijab->sxkjab", Ai, X2)
            Zx = cn("slj,sxkjab->sxklab", Ai, Zx)
            Zx = cn("sac,sxklab->sxklcb", R, Zx)
            Zx = cn("sbd,sxklcb->sxklcd", R, Zx)
            # the antisymmetriser a = 2(1 - P_ab)(1 - P_ij) is self-adjoint and a.a = 8 a, so
            # sum antisym(Zx).Yf = sum Zx.antisym(Yf) = 8 sum Zx.Yf: the completed Z is never formed
            gamma = cn("sxklcd,qklcd->sxq", Zx, Yf, alpha=8.0)
            zxy = cn("sxklcd,qklcd->sxq", Zx, Y2)
        ab = cn("sx,sq->sxq", alpha, beta)
        _axpby(4.0, mixed, 1.0, ab)
        _axpby(1.0, gamma, 1.0, ab)
        _axpby(4.0, ab, 0.0, mixed)          # mixed <- 4 (alpha beta + 4 mixed + gamma)
        # sum x2[ijab] y2[klcd] ia_S_kc jb_S_ld / det(S_oo)^2 (the 8 x2.y2.M_iakc.M_jbld term of aats.py:737) with
        # ia_S_kc = det(S_oo) (P_ai Q_kc + R_ac A_ki):  (U1.P)(W1.Q) + (A U1 R).W1 + (A U2 R).W2 + Z[x2].y2,
        # U1_jb = sum_ia x2 P_ai, U2_ia = sum_jb x2 P_bj, W1_ld = sum_kc y2 Q_kc, W2_kc = sum_ld y2 Q_ld
        U1 = cn("xijab,sai->sxjb", X2, P)
        U2 = cn("xijab,sbj->sxia", X2, P)
        W1 = cn("qklcd,skc->sqld", Y2, Q)
        W2 = cn("qklcd,sld->sqkc", Y2, Q)
        c6 = cn("sx,sq->sxq", cn("sxjb,sbj->sx", U1, P), cn("sqld,sld->sq", W1, Q))
        M1 = cn("sxlb,sbd->sxld", cn("slj,sxjb->sxlb", Ai, U1), R)
        M2 = cn("sxka,sac->sxkc", cn("ski,sxia->sxka", Ai, U2), R)
        _axpby(1.0, cn("sxld,sqld->sxq", M1, W1), 1.0, c6)
        _axpby(1.0, cn("sxkc,sqkc->sxq", M2, W2), 1.0, c6)
        _axpby(1.0, zxy, 1.0, c6)
        # alpha, beta, M, W also give the doubles x singles / singles x doubles tables (see _blocks):
        #   sum_ijab Xf det3(ijab; kc) = -2 alpha Q_kc - 4 M_kc,   sum_klcd Yf det3(ia; klcd) = 2 beta P_ai + 4 N_ia
        return dict(dd=mixed, c6=c6, alpha=alpha, beta=beta, M=M, W=W, Ai=Ai, P=P, Q=Q, R=R)

    def _spatial_terms(self, alpha, beta, normalization):
        m = self.parameters["method"]
        N, N_np, N_nn, N_mp, N_mn = self._spatial_norms(normalization)
        no = self.ndocc
        I = dict.fromkeys(("00", "0D", "D0", "DD", "0S", "S0", "SS", "SD", "DS"), 0)
        R0 = _Tables.get(no, self.nfzc, self.nbf - no).L[0]

        def d2(S):
            return complex(to_host(_det_outer(to_device(np.asarray(S), _C128), no, R0, R0))[0, 0]) ** 2

        a, b = alpha, beta
        if m == "RHF":
            I["00"] = (d2(self.overlap_pp[a][b]) * N_np[a] * N_mp[b] - d2(self.overlap_pn[a][b]) * N_np[a] * N_mn[b]
                       - d2(self.overlap_np[a][b]) * N_nn[a] * N_mp[b] + d2(self.overlap_nn[a][b]) * N_nn[a] * N_mn[b])
            return I
        cisd = m == "CISD"
        if getattr(self, "_rows", None) and alpha not in self._rows:
            self._rows = sorted(set(self._rows) | {int(alpha)})       # a row outside the hint: widen the active set
        amps = self._spatial_amps(normalization)
        A1, A2 = amps[1], amps[2]
        ar = amps["pos"][alpha]                                       # position of alpha in the stacked dR amplitudes
        pick = lambda Am, k: None if Am is None else Am[k]

        def cached(name, S, bra, ket):
            key = ("blk", name, normalization, tuple(self._active_rows()))
            if key not in self._cache:
                self._fill_family(name, normalization, A1, A2)
            if key not in self._cache:
                self._cache[key] = self._block(S, pick(A1, bra), A2[bra], pick(A1, ket), A2[ket])
            return self._cache[key]

        def add(res, sign, ix, iq, s0_N=None, os_N=None, d0=False, od=False):
            I["DD"] += sign * res["DD"][ix, iq]
            if cisd:
                for k in ("SS", "DS", "SD"):
                    I[k] += sign * res[k][ix, iq]
                if s0_N is not None:
                    I["S0"] += sign * res["S0"][ix, iq] * s0_N
                if os_N is not None:
                    I["0S"] += sign * res["0S"][ix, iq] * os_N
                if d0:
                    I["D0"] += sign * res["D0"][ix, iq]
                if od:
                    I["0D"] += sign * res["0D"][ix, iq]

        add(cached("uu", self.overlap_uu, "dR", "dH"), +1, ar, b)
        add(cached(("up", b), self.overlap_up[b], "dR", "t"), +1, ar, 0, s0_N=N_mp[b], d0=True)
        add(cached(("un", b), self.overlap_un[b], "dR", "t"), -1, ar, 0, s0_N=N_mn[b], d0=True)
        add(cached(("pu", a), self.overlap_pu[a], "tc", "dH"), +1, 0, b, os_N=N_np[a], od=True)
        add(cached(("nu", a), self.overlap_nu[a], "tc", "dH"), -1, 0, b, os_N=N_nn[a], od=True)
        key = ("blk4", a, normalization)
        if key not in self._cache:
            # the 12 (beta x pp/pn/np/nn) overlaps of this alpha -- together with those of the next active rows
            # that are still missing, BLK4_GROUP rows per stack: the launches of a stack are latency bound (a
            # few dozen sub-millisecond kernels), so five rows cost little more than one
            todo = [r for r in self._active_rows() if r != a and ("blk4", r, normalization) not in self._cache]
            grp = [a] + todo[:max(0, BLK4_GROUP - 1)]
            stack = [m[r][bb] for r in grp for bb in range(3) for m in (self.overlap_pp, self.overlap_pn,
                                                                          self.overlap_np, self.overlap_nn)]
            res = self._blocks(stack, pick(A1, "tc"), A2["tc"], pick(A1, "t"), A2["t"])
            for j, r in enumerate(grp):
                self._cache[("blk4", r, normalization)] = res[12 * j:12 * j + 12]
        r4 = self._cache[key][4 * b:4 * b + 4]
        # I_00 (aats.py:672-677) from det(S_oo) of the same four overlaps, which their stack has already evaluated
        I["00"] = (r4[0]["dS"] ** 2 * N_np[a] * N_mp[b] - r4[1]["dS"] ** 2 * N_np[a] * N_mn[b]
                   - r4[2]["dS"] ** 2 * N_nn[a] * N_mp[b] + r4[3]["dS"] ** 2 * N_nn[a] * N_mn[b])
        add(r4[0], +1, 0, 0, s0_N=N_mp[b], os_N=N_np[a], d0=True, od=True)
        add(r4[1], -1, 0, 0, s0_N=N_mn[b], os_N=N_np[a], d0=True, od=True)
        add(r4[2], -1, 0, 0, s0_N=N_mp[b], os_N=N_nn[a], d0=True, od=True)
        add(r4[3], +1, 0, 0, s0_N=N_mn[b], os_N=N_nn[a], d0=True, od=True)
        return I

    def prefetch_rows(self, alphas):
        """Hint: the tensor rows (nuclear coordinates alpha) this process is going to ask for.  The
        overlap families that share their amplitude sets -- pu/nu[alpha] for all hinted alpha, up/un[beta]
        for the three beta -- are then evaluated as ONE stack each (all launches batched over the stack)
        instead of one overlap at a time.  Without a hint every row is assumed (single-process use)."""
        self._rows = sorted(set(int(a) for a in alphas))

    def _fill_family(self, name, normalization, A1, A2, max_stack=32):
        if not isinstance(name, tuple) or name[0] not in ("pu", "nu", "up", "un"):
            return
        pick = lambda Am, k: None if Am is None else Am[k]
        if name[0] in ("pu", "nu"):
            rows = getattr(self, "_rows", None) or list(range(len(self.overlap_pu)))
            if name[1] not in rows:
                rows = [name[1]]
            items = [(("pu", a), self.overlap_pu[a]) for a in rows] + [(("nu", a), self.overlap_nu[a]) for a in rows]
            bra, ket = "tc", "dH"
        else:
            items = [(("up", b), self.overlap_up[b]) for b in range(3)] + [(("un", b), self.overlap_un[b]) for b in range(3)]
            bra, ket = "dR", "t"
        rk = tuple(self._active_rows())
        items = [it for it in items if ("blk", it[0], normalization, rk) not in self._cache]
        for i0 in range(0, len(items), max_stack):
            chunk = items[i0:i0 + max_stack]
            res = self._blocks([S for _, S in chunk], pick(A1, bra), A2[bra], pick(A1, ket), A2[ket])
            for (nm, _), r in zip(chunk, res):
                self._cache[("blk", nm, normalization, rk)] = r

    def compute_spatial_aats(self, alpha, beta, normalization="full"):
        """Reference: aats.py:646-1055.  Returns Im(I)/(4 h_R h_B) for one (alpha, beta)."""
        t0 = time.time()
        I = self._spatial_terms(alpha, beta, normalization)
        tot = sum(I.values())
        if config.VERBOSE:
            print(f"AAT element computed in {time.time() - t0} seconds.")
        return (1 / (4 * self.nuc_pert_strength * self.mag_pert_strength)) * np.imag(tot)

    # ---------------------------------------------------------------------------------------
    # spin-orbital brute-force route (aats.py:161-554)
    # ---------------------------------------------------------------------------------------
    def _so_lists(self):
        key = ("solists",)
        if key in self._cache:
            return self._cache[key]
        nocc, nso = 2 * self.ndocc, 2 * self.nbf
        occ, vir = range(nocc), range(nocc, nso)
        tup = {0: np.zeros((1, 0), dtype=np.int32),
               1: np.array([(i, a) for i in occ for a in vir], dtype=np.int32).reshape(-1, 2),
               2: np.array([(i, a, j, b) for i in occ for a in vir for j in occ for b in vir],
                           dtype=np.int32).reshape(-1, 4)}
        L = {}
        for k, t in tup.items():
            out = np.zeros((len(t), nocc), dtype=np.int32)
            check(lib.apyib_so_index_lists(nso, nocc, _i32_host(t)[1] if k else None, len(t), k, _i32_host(out)[1]))
            L[k] = torch.from_numpy(out).to(device())
        self._cache[key] = L
        return L

    def _so_terms(self, alpha, beta, normalization):
        m = self.parameters["method"]
        cisd = m == "CISD_SO"
        nocc = 2 * self.ndocc
        N, N_np, N_nn, N_mp, N_mn = self.compute_normalization(alpha, beta, normalization)
        L = self._so_lists()
        dev = _dev
        if m == "RHF":
            sb = lambda S: spin_block_2_dev(dev(S))                      # aats.py:165-169 (not in place)
        else:
            sb = dev
        a, b = alpha, beta
        Spp, Spn, Snp, Snn = sb(self.overlap_pp[a][b]), sb(self.overlap_pn[a][b]), sb(self.overlap_np[a][b]), sb(self.overlap_nn[a][b])
        one = to_device(np.ones((1, 1)), _C128)

        def bil(S, bk, kk, X, Y):
            """sum_{r,c} X[r] det(S[bra r, ket c]) Y[c]"""
            Z = _det_matvec(S, nocc, L[bk], L[kk], Y.reshape(1, -1))
            return complex(to_host(contract_new("xr,qr->xq", X.reshape(1, -1), Z))[0, 0])

        def stencil(bk, kk, X, Y):
            return (bil(Spp, bk, kk, X, Y) * N_np * N_mp - bil(Spn, bk, kk, X, Y) * N_np * N_mn
                    - bil(Snp, bk, kk, X, Y) * N_nn * N_mp + bil(Snn, bk, kk, X, Y) * N_nn * N_mn)

        I = dict.fromkeys(("00", "0D", "D0", "DD", "0S", "S0", "SS", "SD", "DS"), 0)
        I["00"] = stencil(0, 0, one, one)
        if m == "RHF":
            return I
        U, P_, Ng, Mp, Mn = self.unperturbed_T, self.nuc_pos_T[a], self.nuc_neg_T[a], self.mag_pos_T[b], self.mag_neg_T[b]

        def amps(idx):
            T0 = dev(U[idx])
            tc = _axpby(1.0, T0, 0.0, torch.empty_like(T0), conj_x=True)
            dH = _axpby(-1.0, dev(Mn[idx]), 1.0, dev(Mp[idx]).clone())
            dR = _axpby(1.0, dev(P_[idx]), 0.0, torch.empty_like(T0), conj_x=True)
            dR = _axpby(-1.0, dev(Ng[idx]), 1.0, dR, conj_x=True)
            if idx == 2:   # loop order (i,a,j,b)   <- t2[i][j][a][b]
                T0, tc, dH, dR = (x.permute(0, 2, 1, 3).contiguous() for x in (T0, tc, dH, dR))
            return {0: None, "t": T0, "tc": tc, "dH": dH, "dR": dR}

        A = {1: amps(1) if cisd else None, 2: amps(2)}
        Suu, Sup, Sun = dev(self.overlap_uu), dev(self.overlap_up[b]), dev(self.overlap_un[b])
        Spu, Snu = dev(self.overlap_pu[a]), dev(self.overlap_nu[a])

        def term(bk, kk):
            g = lambda k, name: one if k == 0 else A[k][name]
            pref = (0.25 if bk == 2 else 1.0) * (0.25 if kk == 2 else 1.0)
            tot = stencil(bk, kk, g(bk, "tc"), g(kk, "t"))
            if bk and kk:
                tot += bil(Suu, bk, kk, g(bk, "dR"), g(kk, "dH")) * N * N
            if bk:
                tot += (bil(Sup, bk, kk, g(bk, "dR"), g(kk, "t")) * N * N_mp
                        - bil(Sun, bk, kk, g(bk, "dR"), g(kk, "t")) * N * N_mn)
            if kk:
                tot += (bil(Spu, bk, kk, g(bk, "tc"), g(kk, "dH")) * N_np * N
                        - bil(Snu, bk, kk, g(bk, "tc"), g(kk, "dH")) * N_nn * N)
            return pref * tot

        I["0D"], I["D0"], I["DD"] = term(0, 2), term(2, 0), term(2, 2)
        if cisd:
            I["0S"], I["S0"], I["SS"], I["SD"], I["DS"] = term(0, 1), term(1, 0), term(1, 1), term(1, 2), term(2, 1)
        return I

    def _so_terms_cached(self, alpha, beta, normalization):
        """the nine per-term methods of one element share one evaluation (the reference recomputes per method)"""
        key = ("so_terms", alpha, beta, normalization)
        if key not in self._cache:
            self._cache[key] = self._so_terms(alpha, beta, normalization)
        return self._cache[key]

    def _so_scaled(self, name, alpha, beta, normalization):
        k = 1 / (4 * self.nuc_pert_strength * self.mag_pert_strength)
        return k * np.imag(self._so_terms_cached(alpha, beta, normalization)[name])

    def compute_SO_I_00(self, alpha, beta, normalization): return self._so_scaled("00", alpha, beta, normalization)
    def compute_SO_I_0D(self, alpha, beta, normalization): return self._so_scaled("0D", alpha, beta, normalization)
    def compute_SO_I_D0(self, alpha, beta, normalization): return self._so_scaled("D0", alpha, beta, normalization)
    def compute_SO_I_DD(self, alpha, beta, normalization): return self._so_scaled("DD", alpha, beta, normalization)
    def compute_SO_I_0S(self, alpha, beta, normalization): return self._so_scaled("0S", alpha, beta, normalization)
    def compute_SO_I_S0(self, alpha, beta, normalization): return self._so_scaled("S0", alpha, beta, normalization)
    def compute_SO_I_SS(self, alpha, beta, normalization): return self._so_scaled("SS", alpha, beta, normalization)
    def compute_SO_I_SD(self, alpha, beta, normalization): return self._so_scaled("SD", alpha, beta, normalization)
    def compute_SO_I_DS(self, alpha, beta, normalization): return self._so_scaled("DS", alpha, beta, normalization)

    def compute_SO_aats(self, alpha, beta, normalization="full"):
        """Reference: aats.py:520-554."""
        t0 = time.time()
        I = self._so_terms_cached(alpha, beta, normalization)
        k = 1 / (4 * self.nuc_pert_strength * self.mag_pert_strength)
        tot = sum(k * np.imag(x) for x in I.values())
        if config.VERBOSE:
            print(f"AAT element computed in {time.time() - t0} seconds.")
        return tot
