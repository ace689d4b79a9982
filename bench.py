#!/usr/bin/env python
"""bench.py -- CISD finite-difference AAT time-to-solution per molecule (BASELINE.json metric).

A "step" is one complete pass of the hot path for one molecule: the 6N+7 CISD solves (AO->MO transform, Fock
build, iterations) + MO overlaps + the determinant-overlap assembly of the full (3N,3) AAT tensor.  AO integrals
and the complex-HF SCF of every finite-difference point are host inputs prepared once, outside the timed region
(north star: "stay host-side inputs").

Default workload = BASELINE.json configs[3], the north-star target: (S)-methyloxirane/cc-pVDZ *shape* (nbf = 86,
ndocc = 16, 4 frozen core orbitals, N = 10 atoms, 66+1 finite-difference points, 427 overlap matrices) on synthetic
integrals (no Psi4 in this image).  `--workload h2o2` is configs[1] (H2O2/6-31G shape), `--workload sweep` the
config-5 contraction sweep (spin-orbital CISD iterations, nso = 100 / 200).

  value : seconds per molecule with the per-point AO integrals already resident in HBM and amplitudes kept on the
          device between the solve and the AAT phase;
  e2e   : the same through the public drop-in API with HOST buffers in and out: every point's AO integrals are
          copied host->device inside the timed region (from pinned host memory, on a copy stream that overlaps the
          solves of earlier points) and the tensor / amplitudes come back to the host.

--impl reference : the reference's CPU implementation of the same path (oracle = numpy port of the reference; the
          unmodified reference cannot run this size: it materialises an 8 TB determinant tensor, aats.py:575, and
          needs Psi4).  Every step is a BOUNDED SAMPLE that is really executed (CISD iterations at the workload's
          shape + substituted determinants on all host cores); `ms_per_step` is the wall time of one sample,
          `value` the time-to-solution EXTRAPOLATED from it with the unit counts of SURVEY 8(d) ("extrapolated": true).
          That arm never imports apyib_b200 (host inputs come from the neutral `hostinputs` package).

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload methyloxirane|h2o2|small|sweep]
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE configs[3] (= configs[2]'s molecule): (S)-methyloxirane/cc-pVDZ, frozen core
    "methyloxirane": dict(nbf=86, ndocc=16, natom=10, nfzc=4, algorithm="factorized",
                          label="(S)-methyloxirane/cc-pVDZ shape (nbf=86, ndocc=16, nfzc=4, N=10), CISD FD-AAT, synthetic integrals"),
    # BASELINE configs[1]
    "h2o2": dict(nbf=22, ndocc=9, natom=4, nfzc=0, algorithm="lu",
                 label="H2O2/6-31G shape (nbf=22, ndocc=9, N=4), CISD FD-AAT, synthetic integrals"),
    "small": dict(nbf=10, ndocc=4, natom=2, nfzc=0, algorithm="lu",
                  label="smoke shape (nbf=10, ndocc=4, N=2), CISD FD-AAT, synthetic integrals"),
}
H_R = H_B = 1e-4
METRIC = "cisd_fd_aat_time_to_solution_per_molecule"
UNIT = "s/molecule"
L2_NOTE = "512 MB buffer written between timed steps (flush)"
# FP64 peak on B200 is not in MEASURED_PEAKS.json (bf16 + HBM only); denominator = own DMMA microbenchmark
# (apyib_peak_fp64, re-measured live in every run; profiles/r01_calibration.json: 37.2 TFLOP/s on this pool).
FP64_PEAK_TFLOPS_FALLBACK = 37.2


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def aat_points(natom):
    """The 6N+6 displaced / field points of fin_diff.py:285-370, in the reference's order."""
    pts = [("R", a, +1) for a in range(3 * natom)] + [("R", a, -1) for a in range(3 * natom)]
    return pts + [("B", b, +1) for b in range(3)] + [("B", b, -1) for b in range(3)]


def base_config(wl, method="CISD"):
    return {"workload": wl["label"].replace("CISD", method), "nbf": wl["nbf"], "ndocc": wl["ndocc"], "nfzc": wl["nfzc"], "natom": wl["natom"],
            "fd_points": 6 * wl["natom"] + 7, "overlap_matrices": 7 + 42 * wl["natom"], "h_R": H_R, "h_B": H_B,
            "method": method, "l2": L2_NOTE, "sharding": "fd-points, then tensor rows alpha"}


# ---------------------------------------------------------------------------------------------
# host inputs (untimed): SCF + phase fix for the finite-difference points (numpy only: `hostinputs`)
# ---------------------------------------------------------------------------------------------
def prepare(wl, points=None, method="CISD"):
    import hostinputs as hc
    prov = hc.SyntheticProvider(wl["nbf"], wl["ndocc"], wl["natom"], seed=1000 * 2, nfzc=wl["nfzc"])
    par = {"geom": prov.geometry_string(), "basis": "synthetic", "method": method, "freeze_core": wl["nfzc"] > 0,
           "F_el": [0.0] * 3, "F_mag": [0.0] * 3, "provider": prov, "DIIS": True, "max_iterations": 120,
           "e_convergence": 1e-12, "d_convergence": 1e-12}

    def scf(p):
        w = hc.hf_wfn(hc.Hamiltonian(p))
        w.solve_SCF(p)
        return w

    w0 = scf(par)
    mol = hc.Molecule.from_string(par["geom"])
    geom0 = mol.geometry()
    pts = {}
    for pt in (aat_points(wl["natom"]) if points is None else points):
        p = copy.deepcopy({k: v for k, v in par.items() if k != "provider"})
        p["provider"] = prov
        if pt[0] == "R":
            g = geom0.copy()
            g[pt[1] // 3][pt[1] % 3] += pt[2] * H_R
            mol.set_geometry(g)
            p["geom"] = mol.create_psi4_string_from_molecule()
        else:
            p["F_mag"][pt[1]] += pt[2] * H_B
        w = scf(p)
        S = hc.provider_ao_overlap(w0.H.basis_set, w.H.basis_set)
        d = np.diagonal(w0.C.conj().T @ S @ w.C)
        w.C = w.C * ((d / np.sqrt(d * np.conj(d))) ** -1)[None, :]          # compute_phase, utils.py:427-446
        pts[pt] = w
    return dict(par=par, w0=w0, pts=pts, natom=wl["natom"], prov=prov)


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy restatement of the reference) on a bounded, really executed sample
# ---------------------------------------------------------------------------------------------
def _det_worker(args):
    """One host process: batched np.linalg.det over a slice of the doubly x doubly substituted matrices of one
    overlap (the arithmetic of aats.py:581-618) for `budget` seconds."""
    nbf, no, nf, budget, seed = args
    from oracle import apyib_oracle as orc
    nv = nbf - no
    _, doub = orc.det_index_tables(no, nf, nv)
    S = np.eye(nbf) + 1e-4 * np.random.default_rng(seed).standard_normal((nbf, nbf)).astype(complex)
    rng = np.random.default_rng(seed + 1)
    P = len(doub)
    ncol = min(P, 4096)
    cols = doub[np.sort(rng.choice(P, ncol, replace=False))].reshape(-1, 2, 2)
    nrow = max(1, int(2.0e5 // ncol))
    rows = doub[np.sort(rng.choice(P, min(P, nrow), replace=False))].reshape(-1, 2, 2)
    t0 = time.perf_counter()
    ndet = 0
    while True:
        orc._batched_sub_dets(S, no, rows, cols)
        ndet += len(rows) * len(cols)
        if time.perf_counter() - t0 > budget:
            break
    return ndet, time.perf_counter() - t0


class CpuSampler:
    """Bounded samples of the reference's CPU path at the workload's shape, and the extrapolation to one molecule.

    one sample = (a) `iters` CISD iterations of oracle.solve_CISD (ci_wfn.py:420-574; numpy/BLAS on all cores) on a
    field-free (float64) and on a magnetic-field (complex128) finite-difference point, incl. their AO->MO
    transforms (utils.py:217-279); (b) substituted ndocc x ndocc determinants (aats.py:581-618, batched
    numpy.linalg.det) for `det_budget` seconds in one process per host core (the reference fans the tensor
    elements out over a multiprocessing.Pool, parallel.py:32-42).
    extrapolation (SURVEY 8(d) unit counts): every point costs setup + n_iter iterations, n_iter = iterations of a
    converged oracle solve of the unperturbed point (measured once, untimed); the reference evaluates
    compute_all_dets for 9 overlaps per tensor element (aats.py:714-1008) = 81 N overlap evaluations of
    1 + 2 ov + 2 P + (ov)^2 + 2 P ov + P^2 determinants each."""

    def __init__(self, wl, iters=2, det_budget=2.0):
        from oracle import apyib_oracle as orc
        import multiprocessing as mp
        self.orc, self.wl, self.iters, self.det_budget = orc, wl, iters, det_budget
        self.cores = os.cpu_count() or 1
        try:            # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses all host cores, as it says
            from threadpoolctl import threadpool_limits
            self._blas_limit = threadpool_limits(limits=self.cores)
        except Exception:
            self._blas_limit = None
        self.work = prepare(wl, points=[("B", 0, +1)])
        self.par = self.work["par"]
        self.w_real, self.w_cplx = self.work["w0"], self.work["pts"][("B", 0, +1)]
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.n_iter = None
        # the set-up part of a solve (the constructor: AO->MO transform + Fock build) is timed inside the solve
        self._t_setup = [0.0]
        base, rec = orc._CI, self._t_setup

        class _TimedCI(base):
            def __init__(self, *a, **k):
                t0 = time.perf_counter()
                super().__init__(*a, **k)
                rec[0] = time.perf_counter() - t0

        orc._CI = _TimedCI

    def close(self):
        self.pool.close()

    def converged_iterations(self):
        if self.n_iter is None:
            t0 = time.perf_counter()
            self.n_iter = int(self.orc.solve_CISD(self.par, self.w_real, return_iters=True)[3])
            self.t_converged = time.perf_counter() - t0
        return self.n_iter

    def counts(self):
        wl = self.wl
        o, v, natom = wl["ndocc"] - wl["nfzc"], wl["nbf"] - wl["ndocc"], wl["natom"]
        ov, P = o * v, (o * (o - 1) // 2) * (v * (v - 1) // 2)
        dets = 1 + 2 * ov + 2 * P + ov * ov + 2 * P * ov + P * P
        return dict(points_real=6 * natom + 1, points_complex=6, dets_per_overlap=dets,
                    overlap_evaluations_reference=81 * natom, overlaps_distinct=7 + 42 * natom)

    def sample(self):
        """one really executed sample -> (wall seconds of the sample, measured unit costs)"""
        orc, wl = self.orc, self.wl
        t_start = time.perf_counter()
        p = dict(self.par, max_iterations=self.iters, e_convergence=0.0, d_convergence=0.0)
        unit = {}
        for name, w in (("real", self.w_real), ("complex", self.w_cplx)):
            t0 = time.perf_counter()
            orc.solve_CISD(p, w)                  # set-up (slices + F_MO + ERI_MO, ci_wfn.py:24-47) + `iters` iterations
            t_solve = time.perf_counter() - t0
            unit["setup_" + name] = self._t_setup[0]
            unit["iter_" + name] = max(t_solve - self._t_setup[0], 0.0) / self.iters
        res = self.pool.map(_det_worker, [(wl["nbf"], wl["ndocc"], wl["nfzc"], self.det_budget, k) for k in range(self.cores)])
        unit["dets_timed"] = sum(r[0] for r in res)
        unit["det_aggregate"] = max(r[1] for r in res) / unit["dets_timed"]     # s per determinant, all cores
        return time.perf_counter() - t_start, unit

    def extrapolate(self, unit, n_iter):
        c = self.counts()
        solves = (c["points_real"] * (unit["setup_real"] + n_iter * unit["iter_real"])
                  + c["points_complex"] * (unit["setup_complex"] + n_iter * unit["iter_complex"]))
        dets_ref = c["overlap_evaluations_reference"] * c["dets_per_overlap"] * unit["det_aggregate"]
        dets_distinct = c["overlaps_distinct"] * c["dets_per_overlap"] * unit["det_aggregate"]
        return solves + dets_ref, solves + dets_distinct, solves

    def describe(self, unit, n_iter, wall):
        c, wl = self.counts(), self.wl
        return ("oracle (numpy port of ci_wfn.py:420-574 + aats.py:581-618) on %d host cores, sample of %.1f s really "
                "executed: %d CISD iterations + AO->MO transform of one float64 and one complex128 point at nbf=%d "
                "(setup %.2f / %.2f s, %.3f / %.3f s per iteration) and %d substituted %dx%d determinants in %d "
                "concurrent processes (%.3g us each, aggregate).  EXTRAPOLATED to the molecule with SURVEY 8(d) counts: "
                "%d float64 + %d complex128 points x (setup + %d iterations [converged oracle solve of the unperturbed "
                "point]) + %d overlap evaluations (the reference recomputes compute_all_dets for 9 overlaps per tensor "
                "element; %d distinct overlaps) x %.3g determinants each; the contraction of the 8-index determinant "
                "tensors with the amplitudes is NOT included (lower bound); the unmodified reference cannot run this "
                "shape at all when the tensor exceeds host memory (aats.py:575)"
                % (self.cores, wall, self.iters, wl["nbf"], unit["setup_real"], unit["setup_complex"], unit["iter_real"],
                   unit["iter_complex"], unit["dets_timed"], wl["ndocc"], wl["ndocc"], self.cores,
                   unit["det_aggregate"] * 1e6, c["points_real"], c["points_complex"], n_iter,
                   c["overlap_evaluations_reference"], c["overlaps_distinct"], c["dets_per_overlap"]))


def cpu_baseline_once(wl):
    s = CpuSampler(wl)
    try:
        n_iter = s.converged_iterations()
        wall, unit = s.sample()
        v_ref, v_distinct, v_solves = s.extrapolate(unit, n_iter)
        return {"value": v_ref, "unit": UNIT, "cores": s.cores, "kind": "port", "extrapolated": True,
                "value_distinct_overlaps_only": v_distinct, "value_solves_only": v_solves, "sample_wall_s": wall,
                "sample": s.describe(unit, n_iter, wall)}
    finally:
        s.close()


def reference_arm(args, wl):
    """bench.py --impl reference: K bounded samples really executed (W warm-up samples first), one JSON line."""
    s = CpuSampler(wl)
    try:
        n_iter = s.converged_iterations()
        for _ in range(args.warmup):
            s.sample()
        walls, vals, units = [], [], []
        for _ in range(args.steps):
            wall, unit = s.sample()
            walls.append(wall)
            units.append(unit)
            vals.append(s.extrapolate(unit, n_iter))
        k = int(np.argsort([v[0] for v in vals])[len(vals) // 2])          # the median sample
        v_ref, v_distinct, v_solves = vals[k]
        cores = s.cores
        line = {"impl": "reference", "metric": METRIC, "value": v_ref, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64/c128",
                "data": "synthetic", "config": base_config(wl), "extrapolated": True,
                "note": ("ms_per_step is the wall time of one bounded CPU sample (really executed `steps` times); value "
                         "is the reference's time-to-solution per molecule extrapolated from the median sample"),
                "cpu_baseline": {"value": v_ref, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
                                 "value_distinct_overlaps_only": v_distinct, "value_solves_only": v_solves,
                                 "sample_wall_s": walls[k], "converged_solve_s": s.t_converged,
                                 "sample": s.describe(units[k], n_iter, walls[k])},
                "e2e": {"value": v_ref, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "product_imported": any(m == "apyib_b200" or m.startswith("apyib_b200.") for m in sys.modules)}
        print(json.dumps(line))
    finally:
        s.close()


# ---------------------------------------------------------------------------------------------
# one step of the product path
# ---------------------------------------------------------------------------------------------
def gpu_step(work, rank=0, world=1, dist=None, phases=None):
    import torch
    from apyib_b200.aats import AAT
    from apyib_b200.energy import correlated_many
    from apyib_b200.fin_diff import point_cost
    from apyib_b200.parallel import partition, exchange_points, owned_elements
    par, w0, natom = work["par"], work["w0"], work["natom"]
    n3 = 3 * natom
    mark = (lambda name: phases.append((name, _event()))) if phases is not None else (lambda name: None)
    mark("start")
    # phase 1: the unperturbed point and the 6N+6 displaced / field points, partitioned over the ranks
    allp = [("U", 0, 0)] + aat_points(natom)
    own = partition(allp, [point_cost(p[0]) for p in allp], world)
    my_pts = [p for p, o in zip(allp, own) if o == rank]
    wf = lambda p: w0 if p[0] == "U" else work["pts"][p]
    sols = correlated_many(par, [wf(p) for p in my_pts])          # [(E_corr, [t0, t1, t2])], batched on the device
    mine = {p: T_list for p, (_, T_list) in zip(my_pts, sols)}
    mark("solves")
    import apyib_b200
    if apyib_b200.config.TIMING is not None:             # instrumented step: keep the two phases apart
        work["_timing_solves"], apyib_b200.config.TIMING = apyib_b200.config.TIMING, {}
    if world > 1:
        # exchange step: every AAT element needs T(0), T(R+-alpha), T(B+-beta)  (aats.py:690-711)
        mine = exchange_points(dist, mine, world)
    mark("exchange")
    T = lambda k, i, s: mine[(k, i, s)]
    W = lambda k, i, s: work["pts"][(k, i, s)]
    rows = sorted(set(a for a, _ in owned_elements(n3, rank, world)))
    A = AAT(par, w0, w0.C, w0.H.basis_set, mine[("U", 0, 0)],
            [W("R", a, +1).C for a in range(n3)], [W("R", a, -1).C for a in range(n3)],
            [W("R", a, +1).H.basis_set for a in range(n3)], [W("R", a, -1).H.basis_set for a in range(n3)],
            [T("R", a, +1) for a in range(n3)], [T("R", a, -1) for a in range(n3)],
            [W("B", b, +1).C for b in range(3)], [W("B", b, -1).C for b in range(3)],
            [W("B", b, +1).H.basis_set for b in range(3)], [W("B", b, -1).H.basis_set for b in range(3)],
            [T("B", b, +1) for b in range(3)], [T("B", b, -1) for b in range(3)], H_R, H_B, rows=rows)
    I = np.zeros((n3, 3))
    for a, b in owned_elements(n3, rank, world):
        I[a, b] = A.compute_spatial_aats(a, b)
    mark("aat")
    if world > 1:
        t = torch.from_numpy(I).cuda()
        dist.all_reduce(t)                              # disjoint elements: sum == final gather of the tensor
        I = t.cpu().numpy()
    mark("gather")
    return I


def _event():
    import torch
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def drop_device_caches(work):
    for w in [work["w0"]] + list(work["pts"].values()):
        w.H._apyib_b200_dev = None


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, idx):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.stop_flag = idx, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# roofline of the kernel with the largest share of the step
# ---------------------------------------------------------------------------------------------
def _det_pairs_cmacs(n, nc, k_sub):
    """complex MACs the prefix-shared LU kernel executes per (row list, group) and the group length"""
    gl = nc if k_sub == 1 else nc * (nc - 1) // 2
    npre = n - k_sub
    cmac = (sum((n - 1 - j) * (npre - j) for j in range(npre))          # prefix LU: L column + trailing update
            + k_sub * npre * (npre - 1) // 2 + k_sub * npre               # X = -L21 L11^-1
            + nc * k_sub * npre                                          # candidate columns
            + (gl * 4 + 2 * nc if k_sub == 2 else 2 * gl))               # k x k determinants + table x vector
    return cmac, gl


def roofline_from_timing(timing, wl, step_s, fp64_peak, peak_src):
    """timing: kernel signature -> [(start, end) CUDA events] of ONE full eager step (no graph replay, every launch
    of the contraction / determinant kernels carries events).  Picks the signature with the largest summed time."""
    tot_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in timing.items()}
    if not tot_ms:
        return None
    # launches of the same kernel on the same M x N x K that differ only in the point batch (the chunks of one
    # dtype group) are ONE signature for the ranking; the roofline numbers are those of its largest-batch launches
    fam = lambda k: re.sub(r" b\d+\]$", "]", k)
    fam_ms = {}
    for k, v in tot_ms.items():
        fam_ms[fam(k)] = fam_ms.get(fam(k), 0.0) + v
    top_fam = max(fam_ms, key=fam_ms.get)
    name = max((k for k in tot_ms if fam(k) == top_fam), key=tot_ms.get)
    evs = timing[name]
    avg_ms = tot_ms[name] / len(evs)
    n, nc = wl["ndocc"], wl["nbf"] - wl["ndocc"]
    extra = {}
    m = re.match(r"contract(_tma)?\[(f64|c128) (\d+)x(\d+)x(\d+) b(\d+)\]", name)
    if m:
        M_, N_, K_, nb_ = (int(m.group(i)) for i in (3, 4, 5, 6))
        flops = (8.0 if m.group(2) == "c128" else 2.0) * M_ * N_ * K_ * nb_
        itemsize = 16 if m.group(2) == "c128" else 8
        extra["algorithmic_bytes"] = itemsize * nb_ * (M_ * K_ + K_ * N_ + 2 * M_ * N_)
        kern = ("contract_tma_kernel (FP64 DMMA, TMA-fed) " if m.group(1) else "contract_kernel (FP64 DMMA, LDGSTS gather) ") + name
        tkey = ("contract_tma_kernel " if m.group(1) else "contract_kernel ") + "%s %dx%dx%d" % (m.group(2), M_, N_, K_)
        extra["_nb"] = nb_
        note = ("FP64 tensor (DMMA) roofline; flops = the dense count of the einsum as EXECUTED (2 MNK real / 8 MNK "
                "complex per point; M x N x K and the point batch b are in the kernel name)")
    elif name.startswith("det_pairs"):
        k_sub = int(name.split("k=")[1].split(",")[0])
        nrow, ncol = (int(x) for x in name.split(",")[2].rstrip("]").split("x"))
        n_stack = int(name.split("nS=")[1].rstrip("]")) if "nS=" in name else 1
        cmac, gl = _det_pairs_cmacs(n, nc, k_sub)
        flops = n_stack * nrow * (ncol // gl) * cmac * 8.0
        ndet = nrow * ncol * n_stack
        extra = {"determinants_per_s": ndet / (avg_ms * 1e-3), "executed_complex_macs_per_determinant": cmac / gl,
                 "algorithmic_tflops": ndet * (8.0 / 3.0) * n ** 3 / (avg_ms * 1e-3) / 1e12,
                 "algorithmic_speedup": ndet * (8.0 / 3.0) * n ** 3 / flops}
        kern = "det_pairs_kernel<N=%d,K=%d> (prefix-shared LU + table x vector, %d overlaps per launch) %s" % (n, k_sub, n_stack, name)
        tkey = "det_pairs_kernel"
        note = ("FP64 FMA-pipe roofline on the flops really EXECUTED (one pivoted LU of the n-k unsubstituted columns "
                "per row list and group, one Schur k-vector per candidate column, k x k determinants); the reference's "
                "(8/3)n^3 per determinant (SURVEY 8d U3) is reported as algorithmic_tflops / algorithmic_speedup")
    elif name.startswith("det_matvec") or name.startswith("lemma_matvec"):
        dims = name.split(",")[1].rstrip("]")
        nrow, ncol = (int(x) for x in dims.split("x"))
        n_stack = int(name.split("nS=")[1].rstrip("]")) if "nS=" in name else 1
        ndet = nrow * ncol * n_stack
        lemma = name.startswith("lemma")
        per = 30 * 8.0 if lemma else (8.0 / 3.0) * n ** 3          # ~30 complex mult per <=4x4 cofactor determinant
        flops = ndet * per
        extra = {"determinants_per_s": ndet / (avg_ms * 1e-3),
                 "algorithmic_tflops": ndet * (8.0 / 3.0) * n ** 3 / (avg_ms * 1e-3) / 1e12}
        kern = ("lemma_kernel<fused> " if lemma else "det_tpm_kernel / det_kernel (LU + table x vector) ") + name
        tkey = "lemma_kernel" if lemma else "det_tpm_kernel"
        note = "FP64 FMA-pipe roofline on executed flops (%s per determinant)" % ("~30 complex multiplies" if lemma else "(8/3)n^3")
    else:
        return None
    ach = flops / (avg_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
            "traffic": None, "kernel": kern, "launches_timed": len(evs), "avg_ms": avg_ms,
            "share_of_step": fam_ms[top_fam] * 1e-3 / step_s,
            "launches_of_signature": sum(len(timing[k]) for k in tot_ms if fam(k) == top_fam),
            "note": note + "; timed with CUDA events on the launching stream in ONE extra fully eager step (no CUDA-graph "
                    "replay) right after the timed region; share_of_step = summed time of this signature's launches / "
                    "the step time of the timed region; peak = own FP64 DMMA microbenchmark measured in this run "
                    "(MEASURED_PEAKS.json holds bf16 / HBM only, %s)" % peak_src}
    roof.update(extra)
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tab.get(tkey)
        if ent:
            if "dram_bytes_per_point" in ent:        # contraction captures: one launch of nb_captured points
                roof["traffic"] = ent["dram_bytes_per_point"] * roof.get("_nb", 1)
            else:
                roof["traffic"] = ent.get("dram_bytes_per_launch")
            roof["traffic_source"] = ent.get("source")
    except Exception:
        pass
    roof.pop("_nb", None)
    # the next few signatures, for context (same step)
    top = sorted(tot_ms.items(), key=lambda kv: -kv[1])[:12]
    roof["top_signatures_ms"] = {k: round(v, 2) for k, v in top}
    return roof


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="methyloxirane", choices=sorted(WORKLOADS) + ["sweep"])
    ap.add_argument("--profile-step", action="store_true",
                    help="run the warm-up and ONE device-resident step only (target for ncu launch lists)")
    ap.add_argument("--aat-graph", type=int, default=None, choices=[0, 1],
                    help="replay the AAT overlap stacks from CUDA graphs (default: apyib_b200.config.AAT_USE_GRAPH)")
    ap.add_argument("--aat-algorithm", default=None, choices=["lu", "lemma", "factorized"],
                    help="substituted determinants by batched LU, by the determinant lemma, or in closed form "
                         "(default: per workload -- lu for h2o2, factorized for methyloxirane)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--method", default="CISD", choices=["CISD", "CID", "MP2"],
                    help="correlated method of the finite-difference AAT (BASELINE configs[2] is MP2, configs[3] CISD)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "sweep":
        from tools import bench_sweep
        return bench_sweep.main(args, rank, world, local_rank)
    wl = WORKLOADS[args.workload]
    config = base_config(wl, args.method)

    if args.impl == "reference":
        if args.method != "CISD":
            raise SystemExit("--impl reference samples the CISD workload only")
        if rank == 0:
            reference_arm(args, wl)
        return

    import torch
    import apyib_b200
    from apyib_b200 import _lib, device as dev, fin_diff
    assert fin_diff.aat_points(wl["natom"]) == aat_points(wl["natom"])
    cfg = apyib_b200.config
    cfg.VERBOSE = False
    algo = args.aat_algorithm or wl["algorithm"]
    cfg.AAT_ALGORITHM = algo
    if args.aat_graph is not None:
        cfg.AAT_USE_GRAPH = bool(args.aat_graph)
    use_graph = bool(cfg.AAT_USE_GRAPH)
    extra_cfg = {"aat_algorithm": algo, "aat_graph": use_graph, "pairs_single_vector": bool(cfg.PAIRS_SINGLE_VECTOR)}
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    work = prepare(wl, method=args.method)
    # e2e leg: the host copies of the AO integrals live in pinned (page-locked) memory, registered once here
    pinned = dev.pin_host_inputs([work["w0"]] + list(work["pts"].values()))
    l2buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    per_step, phase_ms = [], []

    def timed_steps(k, device_resident):
        cfg.RETURN_DEVICE = device_resident
        tot, res = 0.0, None
        per_step.clear()
        phase_ms.clear()
        for _ in range(k):
            if not device_resident:
                drop_device_caches(work)
            l2buf.zero_()                                  # 512 MB write, larger than the 126 MB L2
            barrier()
            ph = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            res = gpu_step(work, rank, world, dist, ph)
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            per_step.append(max(e0.elapsed_time(e1) * 1e-3, wall))   # host-orchestrated step: never below wall clock
            tot += per_step[-1]
            phase_ms.append({b[0]: a[1].elapsed_time(b[1]) for a, b in zip(ph[:-1], ph[1:])})
        return tot, res

    def phase_summary():
        keys = list(phase_ms[0]) if phase_ms else []
        return {k + "_ms": round(float(np.mean([p[k] for p in phase_ms])), 3) for k in keys}

    if args.profile_step:
        timed_steps(max(1, args.warmup), True)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()          # ncu --profile-from-start off: only this step is captured
        t, _ = timed_steps(1, True)
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"profile_step_s": t, "launches": _lib.LAUNCHES[0]}))
        return
    sampler = ClockSampler(local_rank)
    # warm-up (also fills the device-resident AO-integral caches and the offset-table cache)
    timed_steps(args.warmup, True)
    if rank == 0:                      # one nvidia-smi poller per job: N of them contend for the driver lock
        sampler.start()
    cfg.TIMING = None
    _lib.LAUNCHES[0] = 0
    t_dev, I_dev = timed_steps(args.steps, True)
    step_times = [round(x, 6) for x in per_step]
    launches = _lib.LAUNCHES[0] // max(args.steps, 1)
    phases_dev = phase_summary()
    # ONE extra fully eager step with events around every contraction / determinant launch (launches replayed from
    # a CUDA graph cannot carry events): which kernel signature has the largest share of the step?
    # (batches one after the other on one stream: a concurrent stream would inflate the per-kernel event times)
    old = (cfg.AAT_USE_GRAPH, cfg.USE_CUDA_GRAPH, cfg.SOLVE_CONCURRENT)
    cfg.AAT_USE_GRAPH, cfg.USE_CUDA_GRAPH, cfg.SOLVE_CONCURRENT, cfg.TIMING, cfg.TIMING_ONLY = False, False, False, {}, None
    timed_steps(1, True)
    eager_step_s = per_step[0]
    timing_aat = cfg.TIMING
    timing = dict(work.pop("_timing_solves", {}))
    for k, v in timing_aat.items():
        timing.setdefault(k, [])
        timing[k] = timing[k] + v
    cfg.TIMING = None
    cfg.AAT_USE_GRAPH, cfg.USE_CUDA_GRAPH, cfg.SOLVE_CONCURRENT = old
    # e2e: host buffers in and out, H2D of every point's AO integrals inside the timed region
    timed_steps(1, False)
    dev.COUNTERS["h2d_bytes"] = dev.COUNTERS["d2h_bytes"] = 0
    t_e2e, I_e2e = timed_steps(args.steps, False)
    h2d, d2h = dev.COUNTERS["h2d_bytes"] // args.steps, dev.COUNTERS["d2h_bytes"] // args.steps
    phases_e2e = phase_summary()
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    if dist is not None:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
        ph = torch.tensor([list(phases_dev.values())], dtype=torch.float64, device="cuda")
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        phases_dev = dict(zip(phases_dev, [round(float(x), 3) for x in ph[0]]))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    assert np.abs(I_dev - I_e2e).max() < 1e-8 * max(1.0, np.abs(I_e2e).max()), "device-resident and e2e legs disagree"

    peaks, peak_src = load_peaks()
    fp64_peak = FP64_PEAK_TFLOPS_FALLBACK
    try:
        import ctypes as C
        fl, ms = C.c_double(), C.c_float()
        _lib.check(_lib.lib.apyib_peak_fp64(1, 4000, C.byref(fl), C.byref(ms)))
        fp64_peak = fl.value / 1e12
    except Exception:
        pass
    roof = roofline_from_timing(timing, wl, t_dev / args.steps, fp64_peak, peak_src)
    if roof is not None:
        roof["eager_step_s"] = eager_step_s
        aat_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in timing_aat.items()}
        roof["top_signatures_aat_phase_ms"] = {k: round(v, 2) for k, v in sorted(aat_ms.items(), key=lambda kv: -kv[1])[:8]}
        roof["timed_kernel_ms_aat_phase"] = round(sum(aat_ms.values()), 2)
        roof["timed_kernel_ms_total"] = round(sum(sum(a.elapsed_time(b) for a, b in v) for v in timing.values()), 2)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.method == "CISD":
        cpu = cpu_baseline_once(wl)
    line = {"metric": METRIC.replace("cisd", args.method.lower()), "value": t_dev / args.steps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64/c128", "data": "synthetic", "config": config,
            "clocks": sampler.result(),
            "e2e": {"value": t_e2e / args.steps, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "host_inputs": "pinned (cudaHostRegister, %d arrays, %.1f GB registered once outside the timed region)"
                                   % (pinned[0], pinned[1] / 1e9), "phases_max_over_ranks": phases_e2e},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "phases_max_over_ranks": phases_dev, "extra": extra_cfg,
            "step_times_s": step_times, "aat_checksum": float(np.abs(I_dev).sum())}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
