"""ctypes binding of libapyib_b200.so (the C-ABI declared in include/apyib_b200.h).

There is NO fallback: if the shared library cannot be loaded the import raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

F64, C128 = 0, 1

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double

# name -> (restype, argtypes); must list every symbol of include/apyib_b200.h
SIGNATURES = {
    "apyib_version": (_int, []),
    "apyib_last_error": (C.c_char_p, []),
    "apyib_device_info": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), _i64p]),
    "apyib_graph_begin": (_int, [_vp]),
    "apyib_graph_end": (_int, [_vp, C.POINTER(_vp)]),
    "apyib_graph_launch": (_int, [_vp, _vp]),
    "apyib_graph_destroy": (_int, [_vp]),
    "apyib_contract": (_int, [_int, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                              _int, _int, _int, _int, _dbl, _dbl, _dbl, _dbl, _int, _i64, _i64, _i64, _vp, _int, _vp, _vp]),
    "apyib_contract_tma": (_int, [_int, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _int, _int,
                                  _dbl, _dbl, _dbl, _dbl, _int, _i64, _i64, _i64, _vp, _vp]),
    "apyib_gather4": (_int, [_int, _vp, _i64p, _int, _vp, _i64p, _i32p, _i64p, _dbl, _i32p, _i64p, _dbl, _vp]),
    "apyib_gather4_batch": (_int, [_int, _vp, _i64p, _int, _int, _i64, _vp, _i64p, _i32p, _i64p, _dbl, _i32p, _i64p, _dbl, _vp]),
    "apyib_gather2": (_int, [_int, _vp, _i64p, _int, _vp, _i64p, _i32p, _i64p, _vp]),
    "apyib_mp2_t2_energy": (_int, [_int, _vp, _i64, _i64, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "apyib_reduce_scratch_len": (_i64, []),
    "apyib_ci_update": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "apyib_symmetrize_ijab": (_int, [_int, _vp, _vp, _i64, _i64, _vp]),
    "apyib_dots": (_int, [_int, _vp, _i64, _int, _vp, _i64, _int, _vp, _vp, _vp]),
    "apyib_diis_push": (_int, [_int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _int, _vp, _vp]),
    "apyib_diis_solve": (_int, [_int, _vp, _int, _int, _vp, _vp, _int, _vp, _vp]),
    "apyib_lincomb_energy_rms": (_int, [_int, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _int, _vp, _vp]),
    "apyib_iter_advance": (_int, [_vp, _vp]),
    "apyib_copy": (_int, [_int, _vp, _vp, _i64, _vp]),
    "apyib_copy_rows": (_int, [_int, _vp, _i64, _vp, _i64, _i64, _int, _dbl, _vp, _vp]),
    "apyib_widen": (_int, [_vp, _vp, _i64, _vp]),
    "apyib_symmetrize_ijab_batch": (_int, [_int, _vp, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "apyib_pack_pairs": (_int, [_int, _vp, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "apyib_unpack_pairs_add": (_int, [_int, _vp, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "apyib_axpby": (_int, [_int, _i64, _dbl, _dbl, _vp, _int, _dbl, _dbl, _vp, _vp]),
    "apyib_det_outer": (_int, [_vp, _int, _int, _vp, _i64, _vp, _i64, _vp, _vp]),
    "apyib_det_matvec": (_int, [_vp, _int, _int, _vp, _i64, _vp, _i64, _vp, _int, _vp, _vp, _vp]),
    "apyib_pack_doubles": (_int, [_vp, _i64, _int, _int, _int, _int, _vp, _i64, _vp, _vp]),
    "apyib_det_matvec_work_len": (_i64, [_i64, _i64, _int, _int]),
    "apyib_det_set_kernel": (_int, [_int]),
    "apyib_det_matvec_pairs_work_len": (_i64, [_i64, _i64, _int, _int, _int, _int, _int]),
    "apyib_det_matvec_pairs": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _int, _vp, _int,
                                      _vp, _vp, _vp]),
    "apyib_det_outer_stack": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "apyib_det_matvec_stack": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp]),
    "apyib_det_matvec_pairs_stack": (_int, [_vp, _int, _int, _int, _int, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _int,
                                            _vp, _i64, _int, _vp, _vp, _vp]),
    "apyib_det_set_pairs_variant": (_int, [_int]),
    "apyib_det_sort_lists": (_int, [_int, _i32p, _i64, _i32p, _f64p, _i32p]),
    "apyib_det_outer_sorted": (_int, [_vp, _int, _int, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "apyib_det_matvec_sorted": (_int, [_vp, _int, _int, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _int, _vp, _vp, _vp]),
    "apyib_lemma_prep_len": (_i64, [_int, _int]),
    "apyib_lemma_prepare": (_int, [_vp, _int, _int, _int, _vp, _vp]),
    "apyib_lemma_outer": (_int, [_vp, _int, _int, _int, _int, _vp, _i64, _int, _vp, _i64, _vp, _vp]),
    "apyib_lemma_matvec_work_len": (_i64, [_i64, _i64, _int, _int]),
    "apyib_lemma_matvec": (_int, [_vp, _int, _int, _int, _int, _vp, _i64, _int, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp]),
    "apyib_get_slices": (_int, [_int, _int, _int, _int, _i32p]),
    "apyib_det_enumeration": (_int, [_int, _int, _int, _i32p, _i64p, _i32p, _i64p]),
    "apyib_det_index_lists": (_int, [_int, _i32p, _i64, _int, _i32p]),
    "apyib_so_index_lists": (_int, [_int, _int, _i32p, _i64, _int, _i32p]),
    "apyib_peak_fp64": (_int, [_int, _int, _f64p, C.POINTER(C.c_float)]),
    "apyib_peak_copy": (_int, [_vp, _vp, _i64, _int, _f64p]),
}


class ApyibB200Error(RuntimeError):
    pass


def _bind(path):
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    return lib


def _load():
    # built in-tree by __graft_entry__.build(); (re)built here when missing or when the sources have changed since
    # (content hash, serialised across ranks).  A library that still lacks a symbol after that fails loudly.
    path = _build.LIB
    if _build._stale():
        path = _build.build_locked()
    return _bind(path)


lib = _load()

# ---- launch accounting (bench.py reports `gpu_launches`) -------------------------------------
# kernels launched per C-ABI call; everything not listed launches nothing on the device
_KERNELS_PER_CALL = {
    "apyib_contract": 1, "apyib_contract_tma": 1, "apyib_gather4": 1, "apyib_gather4_batch": 1, "apyib_gather2": 1, "apyib_mp2_t2_energy": 2, "apyib_ci_update": 1,
    "apyib_symmetrize_ijab": 1, "apyib_dots": 1, "apyib_diis_push": 1, "apyib_diis_solve": 1,
    "apyib_lincomb_energy_rms": 1, "apyib_iter_advance": 1, "apyib_copy": 1, "apyib_copy_rows": 1, "apyib_widen": 1, "apyib_symmetrize_ijab_batch": 1, "apyib_pack_pairs": 1, "apyib_unpack_pairs_add": 1, "apyib_axpby": 1,
    "apyib_det_outer": 1, "apyib_det_matvec": 2, "apyib_det_outer_sorted": 1, "apyib_det_matvec_sorted": 2, "apyib_det_matvec_pairs": 3, "apyib_det_outer_stack": 1, "apyib_det_matvec_stack": 2, "apyib_det_matvec_pairs_stack": 3, "apyib_pack_doubles": 1,
    "apyib_lemma_prepare": 1, "apyib_lemma_outer": 1, "apyib_lemma_matvec": 2,
}
LAUNCHES = [0]


class _Counted:
    __slots__ = ("fn", "n")

    def __init__(self, fn, n):
        self.fn, self.n = fn, n

    def __call__(self, *a):
        LAUNCHES[0] += self.n
        return self.fn(*a)


class _Lib:
    """Thin proxy over the CDLL that counts kernel launches."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name in SIGNATURES:
            fn = getattr(cdll, name)
            n = _KERNELS_PER_CALL.get(name, 0)
            setattr(self, name, _Counted(fn, n) if n else fn)


lib = _Lib(lib)


def check(rc):
    if rc != 0:
        raise ApyibB200Error("libapyib_b200 call failed (%d): %s" % (rc, lib.apyib_last_error().decode()))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise ApyibB200Error("apyib_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
