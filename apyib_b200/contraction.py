"""einsum-style front end of the DMMA contraction kernel (csrc/contract.cu).

Each call is one `opt_einsum.contract(...)` of the reference.  Instead of transposing
operands into GEMM layout, the regrouping of tensor indices into (m | k | n) is encoded in six
integer offset tables (cached per signature on the device); operands may be arbitrary strided
views (slices / swapaxes of a bigger tensor), exactly like the numpy views the reference feeds
to opt_einsum.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import config
from ._lib import lib, check
from .device import dtype_code, ptr, stream_ptr, device

_table_cache = {}


def _offsets(dims, strides):
    """Linearised element offsets of a compound index (last index fastest)."""
    off = np.zeros(1, dtype=np.int64)
    for d, s in zip(dims, strides):
        off = (off[:, None] + (np.arange(d, dtype=np.int64) * s)[None, :]).reshape(-1)
    return off


def _plan(spec, A, B, out):
    key = (spec, tuple(A.shape), tuple(A.stride()), tuple(B.shape), tuple(B.stride()),
           tuple(out.shape), tuple(out.stride()), A.device.index, A.dtype)
    p = _table_cache.get(key)
    if p is not None:
        return p
    ins, o = spec.replace(" ", "").split("->")
    sa, sb = ins.split(",")
    assert len(sa) == A.dim() and len(sb) == B.dim() and len(o) == out.dim(), spec
    size = {}
    for s, t in ((sa, A), (sb, B), (o, out)):
        for ch, d in zip(s, t.shape):
            assert size.setdefault(ch, d) == d, "inconsistent size for index %s in %s" % (ch, spec)
    b_idx = [ch for ch in o if ch in sa and ch in sb]          # batch index (at most one)
    m_idx = [ch for ch in o if ch in sa and ch not in sb]
    n_idx = [ch for ch in o if ch in sb and ch not in sa]
    k_idx = [ch for ch in sa if ch in sb and ch not in o]
    assert len(b_idx) <= 1, "at most one batch index: " + spec
    assert len(m_idx) + len(n_idx) + len(b_idx) == len(o), "repeated output indices unsupported: " + spec
    assert set(sa) == set(m_idx) | set(k_idx) | set(b_idx) and set(sb) == set(n_idx) | set(k_idx) | set(b_idx), \
        "every index must be batch, m, n or k: " + spec
    st = lambda s, t: dict(zip(s, t.stride()))
    stA, stB, stO = st(sa, A), st(sb, B), st(o, out)
    batch = (size[b_idx[0]], stA[b_idx[0]], stB[b_idx[0]], stO[b_idx[0]]) if b_idx else (1, 0, 0, 0)
    # order k by A's layout (largest stride first) so that the fastest k index is contiguous in A if possible
    k_idx.sort(key=lambda ch: -stA[ch])
    dims = lambda idx: [size[ch] for ch in idx]
    tabs = [
        _offsets(dims(m_idx), [stA[ch] for ch in m_idx]),
        _offsets(dims(k_idx), [stA[ch] for ch in k_idx]),
        _offsets(dims(k_idx), [stB[ch] for ch in k_idx]),
        _offsets(dims(n_idx), [stB[ch] for ch in n_idx]),
        _offsets(dims(m_idx), [stO[ch] for ch in m_idx]),
        _offsets(dims(n_idx), [stO[ch] for ch in n_idx]),
    ]
    M, K, N = len(tabs[0]), len(tabs[1]), len(tabs[3])
    fast = lambda stmap, kk, other: int(bool(kk) and (not other or stmap[kk[-1]] <= min(stmap[c] for c in other)))
    a_kfast = fast(stA, k_idx, m_idx)
    b_kfast = fast(stB, k_idx, n_idx)
    dev = torch.from_numpy(np.concatenate(tabs)).to(device())
    sizes = [M, K, K, N, M, N]
    ptrs, off = [], 0
    for s in sizes:
        ptrs.append(dev.data_ptr() + 8 * off)
        off += s
    # plain k-contiguous strided matrices on both sides -> eligible for the TMA-fed kernel
    def linear(tab):
        if len(tab) < 2:
            return int(tab[0]) == 0 and 1
        ld = int(tab[1] - tab[0])
        return ld if ld > 0 and np.array_equal(tab, np.arange(len(tab), dtype=np.int64) * ld) else 0
    tma = None
    # (skinny sides go to the 16-wide tiles of the gather kernel: a 64 x 64 TMA tile would be mostly padding)
    if K >= 16 and M > 16 and N > 16 and linear(tabs[1]) == 1 and linear(tabs[2]) == 1 and tabs[0][0] == 0 and tabs[3][0] == 0:
        lda, ldb = (linear(tabs[0]) if M > 1 else K), (linear(tabs[3]) if N > 1 else K)
        if lda and ldb and lda >= K and ldb >= K and ((M + 63) // 64) * ((N + 63) // 64) * batch[0] >= 148:
            tma = (lda, ldb)
    # split-K for dot-product-like shapes: few output tiles, long K (one CTA would walk K alone).
    # Tile shapes mirror apyib_contract's choice (csrc/contract.cu): 16-wide tiles for skinny sides.
    cplx = A.dtype == torch.complex128
    if M * N <= 16 and K >= 64:
        # dot-product-like (contract_dot_kernel): one CTA of 256 threads per 1 Ki elements of k -- every element
        # costs a chain of two dependent gathers (offset table -> operand), so short per-thread trips hide the
        # latency better than long ones (o^2 v^2 = 13689 at H2O2/6-31G: 34 -> ~12 us per contraction)
        ksplit = int(min(max(1, (K + 1023) // 1024), 128, max(1, (4 * 148) // batch[0]), 65535 // batch[0]))
    else:
        if N <= 16 and M > 16:
            bm, bn = (64 if cplx else 128), 16
        elif M <= 16 and N > 16:
            bm, bn = 16, (64 if cplx else 128)
        else:
            bm = bn = 32
        tiles = ((M + bm - 1) // bm) * ((N + bn - 1) // bn) * batch[0]
        ksplit = 1
        if tiles < 2 * 148 and K >= 2048:
            ksplit = int(min(max(1, (3 * 148) // tiles), (K + 511) // 512, 65535 // batch[0]))
    work = batch[0] * ksplit * M * N if ksplit > 1 else 0          # elements of split-K scratch (per-stream buffer)
    p = (M, N, K, ptrs, a_kfast, b_kfast, dev, batch, ksplit, work, tma)
    _table_cache[key] = p
    return p


_splitk_scratch = {}
SCRATCH_OWNER = [None]        # set by ci_wfn._drive to the batch whose launches are being enqueued (None: caller's stream)


def _work_buffer(dtype, numel):
    """Split-K partial-sum scratch.  Launches of one solve batch are ordered on the batch's stream, so they share one
    buffer; batches that run concurrently on different streams must not.  The buffer is therefore owned by the BATCH
    (SCRATCH_OWNER), not looked up by the current stream: an iteration is captured into a CUDA graph on a separate
    capture stream and replayed on the batch's stream, and the address baked into the graph must be the batch's own.
    Buffers are never freed: captured graphs keep their addresses."""
    if numel == 0:
        return None
    key = (torch.cuda.current_device(), SCRATCH_OWNER[0], dtype)
    bufs = _splitk_scratch.setdefault(key, [])
    if not bufs or bufs[-1].numel() < numel:
        bufs.append(torch.empty(max(numel, 1 << 16), dtype=dtype, device=device()))
    return bufs[-1]


def contract(spec, A, B, out, alpha=1.0, beta=0.0, conj_a=False, conj_b=False):
    """out = alpha * einsum(spec, op(A), op(B)) + beta * out, on the current stream."""
    assert A.dtype == B.dtype == out.dtype, "mixed dtypes: %s %s %s" % (A.dtype, B.dtype, out.dtype)
    M, N, K, ptrs, a_kfast, b_kfast, _keep, batch, ksplit, work, tma = _plan(spec, A, B, out)
    work = _work_buffer(A.dtype, work)
    alpha, beta = complex(alpha), complex(beta)
    if tma is not None and ksplit == 1 and config.USE_TMA:
        with config.timed("contract_tma[%s %dx%dx%d b%d]" % ("c128" if A.dtype == torch.complex128 else "f64", M, N, K, batch[0])):
            rc = lib.apyib_contract_tma(dtype_code(A), ptr(A), ptr(B), ptr(out), M, N, K, tma[0], tma[1],
                                        C.c_void_p(ptrs[4]), C.c_void_p(ptrs[5]), int(conj_a), int(conj_b),
                                        alpha.real, alpha.imag, beta.real, beta.imag,
                                        batch[0], batch[1], batch[2], batch[3], C.c_void_p(0), stream_ptr())
        if rc == 0:
            return out
        if rc != -3:          # APYIB_ERR_UNSUPPORTED -> fall through to the gather kernel
            check(rc)
    if config.TIMING is not None:
        with config.timed("contract[%s %dx%dx%d b%d]" % ("c128" if A.dtype == torch.complex128 else "f64", M, N, K, batch[0])):
            check(_launch(A, B, out, M, N, K, ptrs, a_kfast, b_kfast, conj_a, conj_b, alpha, beta, batch, ksplit, work))
        return out
    check(_launch(A, B, out, M, N, K, ptrs, a_kfast, b_kfast, conj_a, conj_b, alpha, beta, batch, ksplit, work))
    return out


def _launch(A, B, out, M, N, K, ptrs, a_kfast, b_kfast, conj_a, conj_b, alpha, beta, batch, ksplit, work):
    return (lib.apyib_contract(dtype_code(A), ptr(A), ptr(B), ptr(out), M, N, K,
                             *[C.c_void_p(p) for p in ptrs],
                             a_kfast, b_kfast, int(conj_a), int(conj_b),
                             alpha.real, alpha.imag, beta.real, beta.imag,
                             batch[0], batch[1], batch[2], batch[3], C.c_void_p(0), ksplit, ptr(work), stream_ptr()))


def contract_new(spec, A, B, alpha=1.0, conj_a=False, conj_b=False):
    """Allocating variant: returns a fresh dense tensor with the output index order of spec."""
    ins, o = spec.replace(" ", "").split("->")
    sa, sb = ins.split(",")
    size = dict(zip(sa, A.shape))
    size.update(zip(sb, B.shape))
    out = torch.empty([size[ch] for ch in o], dtype=A.dtype, device=A.device)
    return contract(spec, A, B, out, alpha, 0.0, conj_a, conj_b)
