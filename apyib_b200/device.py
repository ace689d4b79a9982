"""Device plumbing: PyTorch owns memory and streams, libapyib_b200 does the arithmetic."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, F64, C128, require_cuda

_scratch = {}
COUNTERS = {"h2d_bytes": 0, "d2h_bytes": 0}     # host<->device traffic of the public API (bench e2e)


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream        # ~0.2 us; current_stream() costs ~10 us
except AttributeError:                                       # pragma: no cover
    _raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream (all libapyib_b200 work is enqueued there)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t):
    if t.dtype == torch.float64:
        return F64
    if t.dtype == torch.complex128:
        return C128
    raise TypeError("apyib_b200 kernels are float64 / complex128 only, got %s" % t.dtype)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_device(x, dtype=None):
    """numpy (or torch) -> contiguous CUDA tensor, float64 or complex128, on the current stream.  A host array that
    lives in pinned memory (see pin_host_inputs) is copied asynchronously; the caller keeps it alive and unchanged
    until the stream has passed the copy (the finite-difference inputs are never mutated on the path)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.ascontiguousarray(x)
        if a.dtype not in (np.float64, np.complex128):
            a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        nbytes = t.numel() * t.element_size()
        COUNTERS["h2d_bytes"] += nbytes
        if nbytes >= _PIN_MIN and t.is_pinned():
            return t.to(device(), non_blocking=True).contiguous()
    return t.to(device(), non_blocking=False).contiguous()


# ---- pinned host inputs ---------------------------------------------------------------------------------------
_PIN_MIN = 1 << 20            # arrays below 1 MiB are not worth a registration
_PINNED = {}                  # address -> (nbytes, array kept alive)


def pin_array(a):
    """Page-lock a C-contiguous numpy array IN PLACE (cudaHostRegister), so that its host->device copies run at full
    PCIe rate and asynchronously.  Returns the number of bytes newly registered (0: already pinned / too small /
    not registrable)."""
    if not isinstance(a, np.ndarray) or not a.flags.c_contiguous or a.nbytes < _PIN_MIN:
        return 0
    addr = a.ctypes.data
    if addr in _PINNED:
        return 0
    require_cuda()
    rc = torch.cuda.cudart().cudaHostRegister(addr, a.nbytes, 0)
    if int(rc) != 0:
        return 0
    _PINNED[addr] = (a.nbytes, a)
    return a.nbytes


def unpin_all():
    for addr in list(_PINNED):
        torch.cuda.cudart().cudaHostUnregister(addr)
        del _PINNED[addr]


def pin_host_inputs(wfns):
    """Register the large host inputs (AO ERIs, T, V, C) of a list of hf_wfn-like objects.  Returns (arrays,
    bytes) newly pinned.  Optional: unpinned inputs work, their copies are synchronous and slower."""
    n = nbytes = 0
    for w in wfns:
        for a in (getattr(w.H, "ERI", None), getattr(w.H, "T", None), getattr(w.H, "V", None), getattr(w, "C", None)):
            b = pin_array(a) if a is not None else 0
            n += 1 if b else 0
            nbytes += b
    return n, nbytes


_copy_streams = {}
_capture_streams = {}


def capture_stream():
    """ONE dedicated stream per device for recording CUDA graphs.  torch.cuda.Stream() hands out streams from a small
    round-robin pool, so a fresh object per capture would sooner or later alias a stream that carries live work."""
    d = torch.cuda.current_device()
    if d not in _capture_streams:
        _capture_streams[d] = torch.cuda.Stream(device=d)
    return _capture_streams[d]



def copy_stream():
    """Dedicated stream for host->device input copies (overlaps the kernels of the compute stream)."""
    d = torch.cuda.current_device()
    if d not in _copy_streams:
        _copy_streams[d] = torch.cuda.Stream(device=d)
    return _copy_streams[d]


def to_host(t, pinned=True):
    """CUDA tensor -> fresh numpy array owned by the caller.  Large results land in page-locked host memory (torch's
    caching pinned allocator), so the copy runs at full PCIe rate and a later re-upload of the same array (the
    amplitudes go back to the device for the AAT assembly) is asynchronous as well."""
    nbytes = t.numel() * t.element_size()
    COUNTERS["d2h_bytes"] += nbytes
    t = t.detach()
    if pinned and t.is_cuda and nbytes >= _PIN_MIN:
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out.numpy()
    return t.cpu().numpy()


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype):
    return torch.zeros(shape, dtype=dtype, device=device())


def reduce_scratch():
    """Zero-initialised partial-sum buffer (+ticket) for the deterministic grid reductions,
    one per (device, stream)."""
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    s = _scratch.get(key)
    if s is None:
        s = torch.zeros(int(lib.apyib_reduce_scratch_len()), dtype=torch.float64, device=device())
        _scratch[key] = s
    return s


def i32(a):
    return (C.c_int32 * len(a))(*[int(v) for v in a])


def i64(a):
    return (C.c_int64 * len(a))(*[int(v) for v in a])


class Graph:
    """One captured iteration (CUDA graph) -- the replacement for re-interpreting the Python
    loop body of the reference's solvers every trip."""

    def __init__(self):
        self._exec = None

    def capture(self, fn):
        """Records the launches `fn` enqueues (nothing executes).  The legacy default stream
        cannot be captured, so recording happens on a private side stream; `fn` must only
        launch libapyib_b200 kernels on the *current* stream and must not allocate."""
        n0 = _lib.LAUNCHES[0]
        side = capture_stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            check(lib.apyib_graph_begin(C.c_void_p(side.cuda_stream)))
            try:
                fn()
            finally:
                h = C.c_void_p()
                rc = lib.apyib_graph_end(C.c_void_p(side.cuda_stream), C.byref(h))
        check(rc)
        self._exec = h
        self._launches = _lib.LAUNCHES[0] - n0
        _lib.LAUNCHES[0] = n0                       # recorded, not executed

    def launch(self):
        _lib.LAUNCHES[0] += self._launches
        check(lib.apyib_graph_launch(self._exec, stream_ptr()))

    def __del__(self):
        try:
            if self._exec is not None:
                lib.apyib_graph_destroy(self._exec)
        except Exception:
            pass
