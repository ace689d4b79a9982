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
    """numpy (or torch) -> contiguous CUDA tensor, float64 or complex128."""
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
        COUNTERS["h2d_bytes"] += t.numel() * t.element_size()
    return t.to(device(), non_blocking=False).contiguous()


def to_host(t):
    COUNTERS["d2h_bytes"] += t.numel() * t.element_size()
    return t.detach().cpu().numpy()


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
        side = torch.cuda.Stream()
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
