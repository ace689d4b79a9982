"""Drop-in counterparts of the hot-path helpers in apyib/utils.py.

Same names, argument meaning and return types (numpy arrays) as the reference; every
function also has a `*_dev` variant returning CUDA tensors, which is what the solver classes use
so that integrals never bounce through the host.  All arithmetic runs in libapyib_b200 kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import lib, check
from .contraction import contract, contract_new
from .device import (to_device, to_host, empty, zeros, dtype_code, ptr, stream_ptr, reduce_scratch, i32, i64, copy_stream)

SPATIAL_METHODS = ("RHF", "MP2", "CID", "CISD")
SO_METHODS = ("MP2_SO", "CID_SO", "CISD_SO")


# ---------------------------------------------------------------------------------------------
# a1  get_slices                                                        (apyib/utils.py:184-213)
# ---------------------------------------------------------------------------------------------
def get_slices(parameters, wfn):
    nfzc = wfn.H.basis_set.n_frozen_core()
    method = parameters["method"]
    if method in SPATIAL_METHODS:
        so = 0
    elif method in SO_METHODS:
        so = 1
    else:
        raise ValueError("unknown method %r" % (method,))
    b = (C.c_int32 * 16)()
    check(lib.apyib_get_slices(int(wfn.nbf), int(wfn.ndocc), int(nfzc), so, b))
    sl = [slice(int(b[2 * k]), int(b[2 * k + 1])) for k in range(8)]
    return sl[:4], sl[4:]


# ---------------------------------------------------------------------------------------------
# upload-once cache of the per-geometry AO integrals (north star: "uploaded once")
# ---------------------------------------------------------------------------------------------
_HOST2DEV = {}           # id(host array) -> (weakref to the array, device copy)


def register_device_copy(host_array, dev_tensor):
    """Remember that `dev_tensor` is the device copy of the numpy array `host_array` (identity, not value): the MO
    coefficients uploaded for the solves are needed again for the MO overlaps of the AAT stage."""
    import weakref
    if isinstance(host_array, np.ndarray):
        if len(_HOST2DEV) > 4096:
            for k in [k for k, (r, _) in _HOST2DEV.items() if r() is None]:
                del _HOST2DEV[k]
        try:
            _HOST2DEV[id(host_array)] = (weakref.ref(host_array), dev_tensor)
        except TypeError:
            pass


def lookup_device_copy(host_array):
    ent = _HOST2DEV.get(id(host_array))
    if ent is not None and ent[0]() is host_array and ent[1].device.index == torch.cuda.current_device():
        return ent[1]
    return None


def _ao_cache(H):
    cache = getattr(H, "_apyib_b200_dev", None)
    if cache is None:
        cache = {}
        try:
            H._apyib_b200_dev = cache
        except AttributeError:
            pass
    return cache


def ao_prefetch(wfns):
    """Start the host->device copies of the AO integrals (h = T + V and the nbf^4 ERIs) of `wfns`, in this order, on
    the copy stream -- the consumers (ao_on_device) make the compute stream wait on a per-point event, so the copies
    of later points overlap the kernels of earlier ones.  Arrays are uploaded in their host dtype (the AO ERIs of a
    magnetic-field point are real: half the bytes of a complex128 copy)."""
    cs = copy_stream()
    todo = [w for w in wfns if not any(k in _ao_cache(w.H) for k in ("raw", "r", "c"))]
    with torch.cuda.stream(cs):
        # ALL small per-point inputs first (h = T + V, MO coefficients, orbital energies): a host->device copy
        # queues behind every copy submitted before it (one DMA engine per direction), so a small synchronous
        # upload issued later -- while 30 GB of AO integrals are in flight -- would stall its stream for the
        # whole upload window (measured: +0.45 s on the first batch at the methyloxirane shape)
        small = [(to_device(np.asarray(w.H.T) + np.asarray(w.H.V)), to_device(np.asarray(w.C)),
                  to_device(np.ascontiguousarray(np.real(np.asarray(w.eps)), dtype=np.float64))) for w in todo]
        for w, (h, Cd, eps) in zip(todo, small):
            cache = _ao_cache(w.H)
            G = cache.get("eri_r")                      # uploaded by the device-assisted SCF already (hostchem)
            if G is None:
                G = to_device(np.asarray(w.H.ERI))
            ev = torch.cuda.Event()
            ev.record(cs)
            cache["raw"] = (h, G, ev)
            cache["C"], cache["eps"], cache["C_host"] = Cd, eps, w.C
            register_device_copy(w.C, Cd)


def wfn_small_on_device(wfn, want_complex):
    """(MO coefficients in the solver's dtype, orbital energies float64) of a point on the device: from the upload
    cache when ao_prefetch / ao_on_device has seen the point (and wfn.C has not been replaced since), else uploaded now."""
    cache = _ao_cache(wfn.H)
    dt = torch.complex128 if want_complex else torch.float64
    Cd = cache.get("C") if cache.get("C_host") is wfn.C else None
    if Cd is None:
        Cd = to_device(np.asarray(wfn.C))
        eps = to_device(np.ascontiguousarray(np.real(np.asarray(wfn.eps)), dtype=np.float64))
        cache["C"], cache["eps"], cache["C_host"] = Cd, eps, wfn.C
        register_device_copy(wfn.C, Cd)
    if Cd.dtype != dt:
        if dt == torch.float64:
            raise TypeError("complex MO coefficients need the complex path")
        Cd = _widen(Cd)
    for x in (Cd, cache["eps"]):
        x.record_stream(torch.cuda.current_stream())
    return Cd, cache["eps"]


def _widen(x):
    out = torch.empty(x.shape, dtype=torch.complex128, device=x.device)
    check(lib.apyib_widen(ptr(out), ptr(x), x.numel(), stream_ptr()))
    return out


def ao_on_device(wfn, want_complex):
    cache = _ao_cache(wfn.H)
    key = "c" if want_complex else "r"
    if key not in cache:
        if "raw" not in cache:
            ao_prefetch([wfn])
        h, G, ev = cache["raw"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for x in (h, G):
            x.record_stream(cur)
        if want_complex:
            cache[key] = tuple(x if x.dtype == torch.complex128 else _widen(x) for x in (h, G))
        else:
            if h.dtype != torch.float64 or G.dtype != torch.float64:
                raise TypeError("complex integrals need the complex path")
            cache[key] = (h, G)
        if cache[key][1] is not G:
            del cache["raw"]                            # the float64 upload of a complex point is not needed again
    return cache[key]


def release_ao(wfn):
    """Drop the device copy of a point's AO integrals (energy-only drivers: thousands of points)."""
    try:
        wfn.H.__dict__.pop("_apyib_b200_dev", None)
    except AttributeError:
        pass


def _is_complex(wfn):
    return bool(np.iscomplexobj(wfn.C) or np.iscomplexobj(wfn.H.T) or np.iscomplexobj(wfn.H.V)
                or np.iscomplexobj(wfn.H.ERI))


# ---------------------------------------------------------------------------------------------
# a2  compute_F_MO                                                      (apyib/utils.py:217-254)
# ---------------------------------------------------------------------------------------------
class LazyScalar:
    """A scalar result that is still on the device (the frozen-core energy of compute_F_MO): reading it back forces a
    host-device synchronisation, so the batched drivers defer it until somebody asks (ci_wfn.E_fc)."""

    def __init__(self, dev, cplx, index=None):
        self.dev, self.cplx, self.index = dev, cplx, index

    def get(self):
        h = to_host(self.dev)
        if self.index is not None:            # one entry of a per-point complex128 vector
            x = h[self.index]
            return complex(x) if self.cplx else float(np.real(x))
        return complex(h[0], h[1]) if self.cplx else float(h[0])


def _axpby_new(alpha, x, beta, y):
    """alpha * x + beta * y as a new tensor (apyib_axpby on a copy of y)"""
    out = y.clone()
    a, b = complex(alpha), complex(beta)
    check(lib.apyib_axpby(dtype_code(out), out.numel(), a.real, a.imag, ptr(x.contiguous()), 0, b.real, b.imag, ptr(out), stream_ptr()))
    return out


def compute_F_MO_dev(parameters, wfn, C_list, lazy=False):
    f, o, v, t = C_list
    cplx = _is_complex(wfn)
    dt = torch.complex128 if cplx else torch.float64
    h, G = ao_on_device(wfn, cplx)
    Cd = wfn_small_on_device(wfn, cplx)[0]
    Gx = G.swapaxes(1, 2)                                   # strided view, no copy

    # (2J - K)[D] for the frozen-core and the active occupied density in ONE pass each over the nbf^4 integrals
    # (utils.py:238-249 contracts them one after the other: four passes)
    fc = parameters["freeze_core"] == True  # noqa: E712 (reference semantics, utils.py:238)
    dens = [contract_new("mp,np->mn", Cd[:, sl], Cd[:, sl], conj_b=True) for sl in ((f, o) if fc else (o,))]
    D = torch.stack(dens)                                           # [x, l, s]
    JK = zeros((len(dens),) + tuple(h.shape), dt)
    contract("xls,mnls->xmn", D, G, JK, alpha=2.0, beta=0.0)        # + 2 J
    contract("xls,mnls->xmn", D, Gx, JK, alpha=-1.0, beta=1.0)      # - K
    E_fc = 0
    if fc:
        h_fc = _axpby_new(1.0, JK[0], 1.0, h)                       # h + (2J - K)[D_fc]
        hs = _axpby_new(1.0, h_fc, 1.0, h)
        e = zeros((2,), torch.float64)
        check(lib.apyib_dots(dtype_code(hs), ptr(dens[0].transpose(0, 1).contiguous()), 0, 1, ptr(hs),
                             hs.numel(), 0, ptr(e), ptr(reduce_scratch()), stream_ptr()))
        E_fc = LazyScalar(e, cplx)
        if not lazy:
            E_fc = E_fc.get()
        h = h_fc
    F_AO = _axpby_new(1.0, JK[-1], 1.0, h)
    Ct = Cd[:, t]
    tmp = contract_new("ij,jq->iq", F_AO, Ct)
    F_MO = contract_new("ip,iq->pq", Ct, tmp, conj_a=True)
    return F_MO, E_fc


def compute_F_MO(parameters, wfn, C_list):
    F, E_fc = compute_F_MO_dev(parameters, wfn, C_list)
    return to_host(F), E_fc


# ---------------------------------------------------------------------------------------------
# a3  compute_ERI_MO                                                    (apyib/utils.py:258-279)
# ---------------------------------------------------------------------------------------------
def compute_ERI_MO_dev(parameters, wfn, C_list):
    """utils.py:258-279: (pq|rs) = sum C*_mp C_nq C*_lr C_gs (mn|lg).  Every quarter transform contracts the LAST
    index of its input and writes its output with the new index FIRST (rotating layouts [m,n,l,g] -> [s,m,n,l] ->
    [r,s,m,n] -> [q,r,s,m] -> [p,q,r,s]), so all four are plain k-contiguous matrix products (nbf^3 x nbf times the
    transposed coefficient block) and run on the TMA-fed DMMA kernel; with the reference's index order three of the
    four contract a strided index and fall back to the element-gather kernel (13 instead of 24 TFLOP/s at nbf = 86)."""
    cplx = _is_complex(wfn)
    dt = torch.complex128 if cplx else torch.float64
    _, G = ao_on_device(wfn, cplx)
    Ct = wfn_small_on_device(wfn, cplx)[0][:, C_list[3]].t().contiguous()        # [t, nbf]: rows = MO, k contiguous
    X = contract_new("mnlg,sg->smnl", G, Ct)
    X = contract_new("smnl,rl->rsmn", X, Ct, conj_b=True)
    X = contract_new("rsmn,qn->qrsm", X, Ct)
    X = contract_new("qrsm,pm->pqrs", X, Ct, conj_b=True)
    return X


def compute_ERI_MO(parameters, wfn, C_list):
    return to_host(compute_ERI_MO_dev(parameters, wfn, C_list))


MO_BATCH_BYTES = 1 << 30        # stack the AO integrals of a group of points only while the stack stays small


def mo_integrals_many(parameters, wfns, C_lists):
    """compute_F_MO + compute_ERI_MO (utils.py:217-279) for a list of points with shared launches: points of
    the same shape and dtype are stacked ([s, ...] leading index = the contraction kernel's batch dimension) --
    9 launches per GROUP instead of 9 per point.  Small molecules only (the 6N+7 points of H2O2/6-31G are
    launch-bound; at cc-pVDZ sizes one point fills the device and stacking would only copy 437 MB per point).
    Returns [(F_MO, E_fc, ERI_MO)] per point; ERI_MO / F_MO of a group are consecutive slices of one tensor."""
    out = [None] * len(wfns)
    groups = {}
    for k, (w, cl) in enumerate(zip(wfns, C_lists)):
        cplx = _is_complex(w)
        key = (cplx, int(w.nbf), tuple((sl.start, sl.stop) for sl in cl))
        groups.setdefault(key, []).append(k)
    for (cplx, nbf, _), idx in groups.items():
        itemsize = 16 if cplx else 8
        if len(idx) == 1 or len(idx) * nbf ** 4 * itemsize > MO_BATCH_BYTES:
            for k in idx:
                F, E_fc = compute_F_MO_dev(parameters, wfns[k], C_lists[k], lazy=True)
                out[k] = (F, E_fc, compute_ERI_MO_dev(parameters, wfns[k], C_lists[k]))
            continue
        dt = torch.complex128 if cplx else torch.float64
        f, o, v, t = C_lists[idx[0]]
        ao = [ao_on_device(wfns[k], cplx) for k in idx]
        h = torch.stack([a[0] for a in ao])
        G = torch.stack([a[1] for a in ao])                               # [s, m, n, l, g]
        npdt = np.complex128 if cplx else np.float64
        Cd = to_device(np.stack([np.asarray(wfns[k].C) for k in idx]).astype(npdt, copy=False), dt)
        Gx = G.swapaxes(2, 3)

        def fock_like(h_in, Cocc):
            D = contract_new("smp,snp->smn", Cocc, Cocc, conj_b=True)
            F = h_in.clone()
            contract("sle,smnle->smn", D, G, F, alpha=2.0, beta=1.0)       # + 2 J
            contract("sle,smnle->smn", D, Gx, F, alpha=-1.0, beta=1.0)     # - K
            return D, F

        E_fc = [0] * len(idx)
        if parameters["freeze_core"] == True:  # noqa: E712 (reference semantics, utils.py:238)
            D_fc, h_fc = fock_like(h, Cd[:, :, f])
            hs = _axpby_new(1.0, h_fc, 1.0, h)                             # h + h_fc (apyib_axpby)
            e = zeros((len(idx),), dt)
            contract("snm,smn->s", D_fc, hs, e, 1.0, 0.0)
            E_fc = [LazyScalar(e, cplx, j) for j in range(len(idx))]
            h = h_fc
        _, F_AO = fock_like(h, Cd[:, :, o])
        Ct = Cd[:, :, t]
        tmp = contract_new("sij,sjq->siq", F_AO, Ct)
        F_MO = contract_new("sip,siq->spq", Ct, tmp, conj_a=True)
        X = contract_new("smnlg,sgx->smnlx", G, Ct)
        X = contract_new("smnlx,slr->smnrx", X, Ct, conj_b=True)
        X = contract_new("snq,smnrx->smqrx", Ct, X)
        X = contract_new("smp,smqrx->spqrx", Ct, X, conj_a=True)
        for j, k in enumerate(idx):
            out[k] = (F_MO[j], E_fc[j], X[j])
    return out


# ---------------------------------------------------------------------------------------------
# a4  spin blocking                                             (apyib/utils.py:283-365, 393-422)
# ---------------------------------------------------------------------------------------------
def spin_block_2_dev(X):
    n0, n1 = X.shape
    Xc = X.contiguous()
    out = empty((2 * n0, 2 * n1), X.dtype)
    check(lib.apyib_gather2(dtype_code(Xc), ptr(Xc), i64([n0, n1]), 1, ptr(out), i64([2 * n0, 2 * n1]),
                            i32([0, 1]), i64([0, 0]), stream_ptr()))
    return out


def gather4(src, spin, out_shape, perm1, start1, c1=1.0, perm2=None, start2=None, c2=0.0, out=None):
    """out[x] = c1*G(start1 + x[perm1]) + c2*G(start2 + x[perm2]) -- see include/apyib_b200.h."""
    assert src.is_contiguous()
    if out is None:
        out = empty(tuple(out_shape), src.dtype)
    assert out.is_contiguous() and tuple(out.shape) == tuple(out_shape)
    p2 = i32(perm2) if perm2 is not None else i32(perm1)
    s2 = i64(start2) if start2 is not None else i64(start1)
    check(lib.apyib_gather4(dtype_code(src), ptr(src), i64(src.shape), int(spin), ptr(out), i64(out_shape),
                            i32(perm1), i64(start1), float(c1), p2, s2, float(c2), stream_ptr()))
    return out


def gather4_stack(srcs, spin, out_shape, perm1, start1, c1=1.0, perm2=None, start2=None, c2=0.0):
    """gather4 of the same block from every tensor of `srcs` -> out[len(srcs), *out_shape].  When the sources are
    consecutive slices of one stacked tensor (what mo_integrals_many produces) this is ONE launch (grid.y = point),
    otherwise one launch per source."""
    nb = len(srcs)
    out = empty((nb,) + tuple(out_shape), srcs[0].dtype)
    step = srcs[0].numel() * srcs[0].element_size()
    base = srcs[0].data_ptr()
    if nb > 1 and all(x.is_contiguous() and x.shape == srcs[0].shape and x.data_ptr() == base + k * step
                      for k, x in enumerate(srcs)):
        p2 = i32(perm2) if perm2 is not None else i32(perm1)
        s2 = i64(start2) if start2 is not None else i64(start1)
        check(lib.apyib_gather4_batch(dtype_code(srcs[0]), ptr(srcs[0]), i64(srcs[0].shape), int(spin), nb, srcs[0].numel(),
                                      ptr(out), i64(out_shape), i32(perm1), i64(start1), float(c1), p2, s2, float(c2),
                                      stream_ptr()))
        return out
    for k, x in enumerate(srcs):
        gather4(x, spin, out_shape, perm1, start1, c1, perm2, start2, c2, out=out[k])
    return out


def spin_block_4_dev(X):
    n = list(X.shape)
    return gather4(X.contiguous(), 1, [2 * d for d in n], [0, 1, 2, 3], [0, 0, 0, 0])


def compute_F_SO(wfn, F_MO):
    return to_host(spin_block_2_dev(to_device(F_MO)))


def compute_ERI_SO(wfn, ERI_MO):
    return to_host(spin_block_4_dev(to_device(ERI_MO)))


def compute_so_overlap(nbf, mo_overlap):
    return to_host(spin_block_2_dev(to_device(mo_overlap)))


# ---------------------------------------------------------------------------------------------
# a14  MO overlap                                                       (apyib/utils.py:370-388)
# ---------------------------------------------------------------------------------------------
def mo_overlap_dev(C_bra, S_ao, C_ket):
    """C_bra^H S_AO C_ket on the device; S_AO is the (mixed-geometry) AO overlap, a host input."""
    cplx = any(np.iscomplexobj(x) for x in (C_bra, S_ao, C_ket))
    dt = torch.complex128 if cplx else torch.float64
    Cb, S, Ck = to_device(C_bra, dt), to_device(S_ao, dt), to_device(C_ket, dt)
    tmp = contract_new("mn,nq->mq", S, Ck)
    return contract_new("mp,mq->pq", Cb, tmp, conj_a=True)


def mo_overlaps_dev(triples, host=False):
    """[C_bra^H S_AO C_ket for (C_bra, S_AO, C_ket) in triples] with TWO batched contraction launches per
    dtype group (real / complex) instead of two per matrix: the finite-difference AAT needs 1 + 6 + 6N + 36N
    overlaps of identical shape (aats.py:53-115) built from only 6N + 7 distinct coefficient matrices and 6N + 1
    distinct AO overlaps.  Every DISTINCT operand (by identity) goes to the device once -- coefficient matrices
    that were uploaded for the solves (ao_prefetch) or came through the NCCL exchange are not uploaded at all --
    and the per-overlap operand stacks are gathered on the device.  Returns device tensors, float64 where all
    three inputs are real (like numpy would), complex128 otherwise; host=True returns numpy arrays instead (one
    copy per group)."""
    out = [None] * len(triples)
    uniq, ops = {}, []                                   # id(operand) -> position, [(device tensor, is complex)]

    def slot(x):
        k = uniq.get(id(x))
        if k is None:
            if isinstance(x, torch.Tensor):
                d = x if x.is_cuda else to_device(x)
            else:
                d = lookup_device_copy(x)
                if d is None:
                    d = to_device(np.asarray(x))
            k = uniq[id(x)] = len(ops)
            ops.append(d)
        return k

    idx = [tuple(slot(x) for x in t) for t in triples]
    groups = {}
    for k, t in enumerate(idx):
        cplx = any(ops[j].dtype == torch.complex128 for j in t)
        groups.setdefault((cplx,) + tuple(tuple(ops[j].shape) for j in t), []).append(k)
    for key, members in groups.items():
        dt = torch.complex128 if key[0] else torch.float64
        used = sorted(set(j for k in members for j in idx[k]))
        pos = {j: p for p, j in enumerate(used)}
        shapes = set(tuple(ops[j].shape) for j in used)
        if len(shapes) == 1:                             # square case: one stack of all distinct operands
            U = torch.stack([ops[j] if ops[j].dtype == dt else _widen(ops[j]) for j in used])
            sel = lambda c: U.index_select(0, torch.tensor([pos[idx[k][c]] for k in members], device=U.device))
            Cb, S, Ck = sel(0), sel(1), sel(2)
        else:
            cast = lambda j: ops[j] if ops[j].dtype == dt else _widen(ops[j])
            Cb, S, Ck = (torch.stack([cast(idx[k][c]) for k in members]) for c in range(3))
        tmp = contract_new("smn,snq->smq", S, Ck)
        res = contract_new("smp,smq->spq", Cb, tmp, conj_a=True)
        if host:
            # ONE device->host copy per dtype group, into ordinary memory: the overlaps live as long as the AAT
            # object, and a fresh page-locked block per object costs more (cudaHostAlloc) than the copy saves
            res = to_host(res, pinned=False)
        for j, k in enumerate(members):
            out[k] = res[j]
    return out


def compute_mo_overlap(ndocc, nbf, bra_basis, bra_wfn, ket_basis, ket_wfn, ao_overlap=None):
    """Reference signature plus an explicit AO overlap: Psi4's mixed-basis `ao_overlap` is a host
    input (the integral provider supplies it)."""
    if ao_overlap is None:
        from .hostchem import provider_ao_overlap
        ao_overlap = provider_ao_overlap(bra_basis, ket_basis)
    return to_host(mo_overlap_dev(bra_wfn, ao_overlap, ket_wfn))


# ---------------------------------------------------------------------------------------------
# a11  compute_phase                                                    (apyib/utils.py:427-446)
# ---------------------------------------------------------------------------------------------
def compute_phase(ndocc, nbf, unperturbed_basis, unperturbed_wfn, ket_basis, ket_wfn, ao_overlap=None):
    S = compute_mo_overlap(ndocc, nbf, unperturbed_basis, unperturbed_wfn, ket_basis, ket_wfn, ao_overlap)
    d = np.diagonal(S)
    N = np.sqrt(d * np.conjugate(d))
    phase = d / N
    return np.asarray(ket_wfn) * (phase ** -1)[None, :]


# ---------------------------------------------------------------------------------------------
# a10  solve_general_DIIS                                               (apyib/utils.py:104-140)
# ---------------------------------------------------------------------------------------------
def solve_general_DIIS(parameters, res_vec, t_vec, e_iter, t_iter, iteration, min_DIIS=1, max_DIIS=7):
    """The reference's DIIS step with its own calling convention (numpy arrays in and out: error / amplitude history
    as columns, returns (t_vec, e_iter, t_iter)).  The CI solvers of this package keep their history in device ring
    buffers and never call this; it exists for external callers of apyib.utils and runs the same kernels: Gram
    matrix B = e^H e by the contraction kernel, the bordered (m+1) x (m+1) solve by `apyib_diis_solve`, the
    extrapolation t = sum_j c_j t_j by the contraction kernel."""
    e_iter, t_iter = np.asarray(e_iter), np.asarray(t_iter)
    while e_iter.shape[1] > max_DIIS:                                     # utils.py:109-112
        e_iter, t_iter = e_iter[:, 1:], t_iter[:, 1:]
    if iteration != 1:                                                    # utils.py:117-119
        e_iter = np.hstack((e_iter, np.atleast_2d(np.asarray(res_vec)).T))
        t_iter = np.hstack((t_iter, np.atleast_2d(np.asarray(t_vec)).T))
    m = e_iter.shape[1]
    if m > 8:
        raise ValueError("solve_general_DIIS: at most 8 history vectors (max_DIIS <= 7)")
    cplx = np.iscomplexobj(e_iter) or np.iscomplexobj(t_iter)
    E = to_device(np.ascontiguousarray(e_iter.T).astype(np.complex128), torch.complex128)      # [m, len]
    T = to_device(np.ascontiguousarray(t_iter.T).astype(np.complex128), torch.complex128)
    B = contract_new("mo,no->mn", E, E, conj_a=True)                      # utils.py:121
    c = zeros((16,), torch.float64)
    check(lib.apyib_diis_solve(1, ptr(B), m, m, C.c_void_p(0), ptr(c), 1, C.c_void_p(0), stream_ptr()))   # :122-135
    cc = torch.view_as_complex(c.view(8, 2))[:m].contiguous()
    t_new = to_host(contract_new("m,mo->o", cc, T))                       # utils.py:138
    return (t_new if cplx else np.ascontiguousarray(t_new.real)), e_iter, t_iter


# ---------------------------------------------------------------------------------------------
# a21  perturbed MO integrals of the analytic route              (apyib/analytic_aats.py:730-733, 977-981)
# ---------------------------------------------------------------------------------------------
def build_dERI(U, ERI_full, nfzc, kind, core=None):
    """dERI_dH (kind "H", analytic_aats.py:730-733) / dERI_dR (kind "R", :977-981): four nbf^5 contractions of the CPHF
    coefficients U (nbf, nbf) with the PHYSICISTS' MO integrals over all nbf orbitals (`ERI` of the reference at that
    point, analytic_aats.py:27-33), t = [nfzc, nbf); the bra terms carry a minus sign for the magnetic field.  `core`
    (kind "R") is the derivative-integral term `ERI_core[a]`, a host input.  Returns a numpy array (n_t^4)."""
    cplx = any(np.iscomplexobj(x) for x in (U, ERI_full) + (() if core is None else (core,)))
    dt = torch.complex128 if cplx else torch.float64
    Ud, W = to_device(np.asarray(U), dt), to_device(np.asarray(ERI_full), dt)
    nbf = Ud.shape[0]
    t = slice(int(nfzc), nbf)
    nt = nbf - int(nfzc)
    Ut = Ud[:, t]
    sgn = -1.0 if kind == "H" else 1.0
    out = zeros((nt, nt, nt, nt), dt) if core is None else to_device(np.asarray(core), dt).clone()
    contract("tr,pqts->pqrs", Ut, W[t, t, :, t], out, 1.0, 1.0)
    contract("ts,pqrt->pqrs", Ut, W[t, t, t, :], out, 1.0, 1.0)
    contract("tp,tqrs->pqrs", Ut, W[:, t, t, t], out, sgn, 1.0)
    contract("tq,ptrs->pqrs", Ut, W[t, :, t, t], out, sgn, 1.0)
    return to_host(out)
