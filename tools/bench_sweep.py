"""bench.py --workload sweep: BASELINE.json configs[4], the synthetic complex spin-orbital CISD sigma sweep
(SURVEY 8(d) "Config 5 sweep"): nso in {100, 200[, 300]}, O = nso/5, V = nso - O, complex128, full
ci_wfn.solve_CISD_SO iterations (ci_wfn.py:263-416) with DIIS on and the convergence test disabled.

Per nso the time of ONE iteration (residual build + update + DIIS + energy/rms) is the difference between a
(n0 + n)-iteration and an n0-iteration solve divided by n, so the one-off set-up (AO->MO transform, spin-blocked
antisymmetrised integral blocks) is excluded; flops = 8 W1(O, V) (SURVEY 8(d) unit U1: the dense count of the
reference's einsums), reported against the FP64 DMMA peak measured in the same run.
nso = 400 needs packed a<b, c<d storage (the dense <ab||cd> block is 168 GB) and is not run.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def W1(O, V):
    return (O ** 2 * V ** 4 + 4 * O ** 3 * V ** 3 + O ** 4 * V ** 2 + 4 * O ** 2 * V ** 3 + 4 * O ** 3 * V ** 2
            + (O ** 2 * V ** 3 + O ** 3 * V ** 2 + 2 * O ** 2 * V ** 2 + O ** 2 * V + O * V ** 2))


class _Basis:
    def n_frozen_core(self):
        return 0


class _H:
    pass


class _Wfn:
    pass


def synthetic_point(nbf, ndocc, seed):
    """SURVEY 8(d) generator: Hermitian-symmetric complex (pq|rs), eps with a 4 Eh gap, C = 1."""
    rng = np.random.default_rng(seed)
    g = (0.25 / nbf) * rng.standard_normal((nbf,) * 4)
    g = g + 0.1j * (0.25 / nbf) * rng.standard_normal((nbf,) * 4)
    g = g + g.transpose(2, 3, 0, 1)
    g = g + g.transpose(1, 0, 3, 2).conj()
    eps = np.sort(rng.standard_normal(nbf))
    eps[ndocc:] += 4.0
    w, h = _Wfn(), _H()
    h.T, h.V, h.ERI, h.E_nuc, h.basis_set = np.diag(eps).astype(complex), np.zeros((nbf, nbf), dtype=complex), g, 0.0, _Basis()
    w.C, w.eps, w.nbf, w.ndocc, w.E_SCF, w.H = np.eye(nbf, dtype=complex), eps, nbf, ndocc, 0.0, h
    return w


def main(args, rank, world, local_rank):
    if rank != 0:
        return
    import ctypes as C
    import torch
    import apyib_b200
    from apyib_b200 import _lib
    torch.cuda.set_device(local_rank)
    apyib_b200.config.VERBOSE = False
    apyib_b200.config.RETURN_DEVICE = True
    fl, ms = C.c_double(), C.c_float()
    _lib.check(_lib.lib.apyib_peak_fp64(1, 4000, C.byref(fl), C.byref(ms)))
    peak = fl.value / 1e12
    sizes = [int(x) for x in os.environ.get("APYIB_SWEEP_NSO", "100,200").split(",")]
    n0, n = 2, max(args.steps, 1)
    rows = []
    for nso in sizes:
        nbf, no = nso // 2, nso // 10
        O, V = 2 * no, nso - 2 * no
        w = synthetic_point(nbf, no, 5000 + nso)
        par = {"method": "CISD_SO", "freeze_core": False, "DIIS": True, "e_convergence": 0.0, "d_convergence": 0.0}
        times = {}
        for its in (n0, n0, n0 + n):                      # first pass = warm-up (offset tables, module load)
            ci = apyib_b200.ci_wfn(dict(par, max_iterations=its), w)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            E = ci.solve_CISD_SO()[0]
            torch.cuda.synchronize()
            times[its] = time.perf_counter() - t0
            del ci
        t_iter = (times[n0 + n] - times[n0]) / n
        flops = 8.0 * W1(O, V)
        rows.append({"nso": nso, "O": O, "V": V, "ms_per_iteration": 1e3 * t_iter, "U1_flops": flops,
                     "tflops": flops / t_iter / 1e12, "frac_of_fp64_dmma_peak": flops / t_iter / 1e12 / peak,
                     "setup_plus_%d_iterations_s" % n0: times[n0], "E_corr_re": float(np.real(E))})
        w.H._apyib_b200_dev = None
        del w
        torch.cuda.empty_cache()
    last = rows[-1]
    print(json.dumps({"metric": "cisd_so_iteration_fp64_tflops", "value": last["tflops"], "unit": "TFLOP/s", "n_gpus": 1,
                      "steps": n, "warmup": n0, "ms_per_step": last["ms_per_iteration"], "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                      "config": {"workload": "synthetic complex spin-orbital CISD sigma sweep (BASELINE configs[4]), "
                                             "full solve_CISD_SO iterations, nso = %s" % sizes, "nso": sizes},
                      "roofline": {"bound": "tensor", "achieved": last["tflops"], "peak": peak, "unit": "TFLOP/s",
                                   "frac": last["tflops"] / peak, "traffic": None,
                                   "note": "whole iteration (all ~20 contractions + streaming kernels), U1 = 8 W1(O,V) flops"},
                      "sweep": rows}))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    main(ap.parse_args(), 0, 1, 0)
