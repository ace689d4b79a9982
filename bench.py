#!/usr/bin/env python
"""bench.py -- CISD finite-difference AAT time-to-solution per molecule (BASELINE.json metric).

A "step" is one complete pass of the hot path for one molecule: the 6N+7 CISD solves
(AO->MO transform, Fock build, iterations) + MO overlaps + the determinant-overlap assembly of
the full (3N,3) AAT tensor.  AO integrals and the complex-HF SCF of every finite-difference point
are host inputs prepared once, outside the timed region (north star: "stay host-side inputs").

Workload at N=1 (and for every N): BASELINE.json configs[1], H2O2/6-31G *shape* -- nbf=22,
ndocc=9, 4 atoms, CISD, 30+1 points -- on synthetic integrals (no Psi4 in this image).

  value : seconds per molecule with the per-point AO integrals already resident in HBM and
          amplitudes kept on the device between the solve and the AAT phase;
  e2e   : the same through the public drop-in API with host (numpy) buffers in and out:
          ci_wfn(parameters, wfn).solve_CISD(), AAT(...), compute_spatial_aats(alpha, beta).

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload h2o2|small]
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: nbf, ndocc, natom, nfzc
    "h2o2": dict(nbf=22, ndocc=9, natom=4, nfzc=0, label="H2O2/6-31G shape (nbf=22, ndocc=9, N=4), CISD FD-AAT, synthetic integrals"),
    # BASELINE configs[2-3] shape: (S)-methyloxirane/cc-pVDZ, frozen core; only feasible with the factorised AAT
    "methyloxirane": dict(nbf=86, ndocc=16, natom=10, nfzc=4, label="(S)-methyloxirane/cc-pVDZ shape (nbf=86, ndocc=16, nfzc=4, N=10), CISD FD-AAT, synthetic integrals"),
    "small": dict(nbf=10, ndocc=4, natom=2, nfzc=0, label="smoke shape (nbf=10, ndocc=4, N=2), CISD FD-AAT, synthetic integrals"),
}
H_R = H_B = 1e-4
METRIC = "cisd_fd_aat_time_to_solution_per_molecule"
UNIT = "s/molecule"


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# FP64 peak on B200 is not in MEASURED_PEAKS.json (bf16 + HBM only); denominator = own DMMA
# microbenchmark measured on this pool (profiles/r01_calibration.json), re-measured live when possible.
FP64_PEAK_TFLOPS_FALLBACK = 37.2


# ---------------------------------------------------------------------------------------------
# host inputs (untimed): SCF + phase fix for every finite-difference point
# ---------------------------------------------------------------------------------------------
def prepare(wl):
    from apyib_b200 import hostchem as hc
    from apyib_b200.fin_diff import aat_points
    prov = hc.SyntheticProvider(wl["nbf"], wl["ndocc"], wl["natom"], seed=1000 * 2, nfzc=wl["nfzc"])
    par = {"geom": prov.geometry_string(), "basis": "synthetic", "method": "CISD", "freeze_core": wl["nfzc"] > 0,
           "F_el": [0.0] * 3, "F_mag": [0.0] * 3, "provider": prov, "DIIS": True, "max_iterations": 120,
           "e_convergence": 1e-12, "d_convergence": 1e-12}

    def scf(p):
        H = hc.Hamiltonian(p)
        w = hc.hf_wfn(H)
        w.solve_SCF(p)
        return w

    w0 = scf(par)
    mol = hc.Molecule.from_string(par["geom"])
    geom0 = mol.geometry()
    pts = {}
    for pt in aat_points(wl["natom"]):
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


def work_items(work, rank, world):
    from apyib_b200.fin_diff import aat_points, point_cost
    from apyib_b200.parallel import partition
    pts = aat_points(work["natom"])
    own = partition(pts, [point_cost(p[0]) for p in pts], world)
    return pts, own


# ---------------------------------------------------------------------------------------------
# one step of the product path
# ---------------------------------------------------------------------------------------------
def gpu_step(work, rank=0, world=1, dist=None):
    import torch
    import apyib_b200
    from apyib_b200.aats import AAT
    par, w0, natom = work["par"], work["w0"], work["natom"]
    pts, own = work_items(work, rank, world)
    n3 = 3 * natom

    from apyib_b200.ci_wfn import solve_many
    my_pts = [p for p, o in zip(pts, own) if o == rank]
    # every rank needs the unperturbed amplitudes; all points of the rank are solved together
    sols = solve_many("CISD", par, [w0] + [work["pts"][p] for p in my_pts])
    T0 = [1, sols[0][1], sols[0][2]]
    mine = {p: [1, r[1], r[2]] for p, r in zip(my_pts, sols[1:])}
    if world > 1:
        # exchange step: every AAT element needs T(R+-alpha), T(B+-beta)  (aats.py:690-711)
        from apyib_b200.parallel import exchange_points
        mine = exchange_points(dist, mine, world)
    T = lambda k, i, s: mine[(k, i, s)]
    W = lambda k, i, s: work["pts"][(k, i, s)]
    A = AAT(par, w0, w0.C, w0.H.basis_set, T0,
            [W("R", a, +1).C for a in range(n3)], [W("R", a, -1).C for a in range(n3)],
            [W("R", a, +1).H.basis_set for a in range(n3)], [W("R", a, -1).H.basis_set for a in range(n3)],
            [T("R", a, +1) for a in range(n3)], [T("R", a, -1) for a in range(n3)],
            [W("B", b, +1).C for b in range(3)], [W("B", b, -1).C for b in range(3)],
            [W("B", b, +1).H.basis_set for b in range(3)], [W("B", b, -1).H.basis_set for b in range(3)],
            [T("B", b, +1) for b in range(3)], [T("B", b, -1) for b in range(3)], H_R, H_B)
    from apyib_b200.parallel import owned_elements
    I = np.zeros((n3, 3))
    A.prefetch_rows([a for a, _ in owned_elements(n3, rank, world)])
    for a, b in owned_elements(n3, rank, world):
        I[a, b] = A.compute_spatial_aats(a, b)
    if world > 1:
        t = torch.from_numpy(I).cuda()
        dist.all_reduce(t)                              # disjoint elements: sum == final gather of the tensor
        I = t.cpu().numpy()
    return I


def drop_device_caches(work):
    for w in [work["w0"]] + list(work["pts"].values()):
        w.H._apyib_b200_dev = None


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle (numpy restatement of the reference) on a bounded sample
# ---------------------------------------------------------------------------------------------
def _det_worker(args):
    """One host process of the CPU baseline: batched np.linalg.det over a slice of the substituted
    matrices of one overlap (the arithmetic of aats.py:581-618), for `budget` seconds."""
    nbf, no, nf, budget, seed = args
    from oracle import apyib_oracle as orc
    nv = nbf - no
    sing, doub = orc.det_index_tables(no, nf, nv)
    S = np.eye(nbf) + 1e-4 * np.random.default_rng(seed).standard_normal((nbf, nbf)).astype(complex)
    d_sub = doub.reshape(-1, 2, 2)
    P = len(d_sub)
    nrow = max(1, min(P, int(2.0e5 // max(P, 1)) or 1))
    t0 = time.perf_counter()
    ndet = 0
    while True:
        orc._batched_sub_dets(S, no, d_sub[:nrow], d_sub)
        ndet += nrow * P
        if time.perf_counter() - t0 > budget:
            break
    return ndet, time.perf_counter() - t0


def cpu_sample(work, budget_s=20.0):
    """CPU baseline on ALL host cores: (i) oracle.solve_CISD on finite-difference points (numpy/BLAS
    threads) and (ii) the substituted-determinant evaluation of compute_all_dets (aats.py:581-618,
    batched np.linalg.det) on a slice of one overlap, run concurrently in one process per core (the
    reference fans elements out over a multiprocessing.Pool, parallel.py:32-42); both are scaled to the
    whole molecule by the unit counts of SURVEY 8(d).  Returns (seconds_per_molecule, description)."""
    import multiprocessing as mp
    from oracle import apyib_oracle as orc
    par, natom = work["par"], work["natom"]
    w0 = work["w0"]
    npts = 6 * natom + 7
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    nsolve = 0
    for w in [w0] + list(work["pts"].values())[:3] + list(work["pts"].values())[-1:]:
        orc.solve_CISD(par, w)
        nsolve += 1
        if time.perf_counter() - t0 > budget_s / 3:
            break
    t_solve = (time.perf_counter() - t0) / nsolve
    no, nv = w0.ndocc, w0.nbf - w0.ndocc
    sing, doub = orc.det_index_tables(no, par_nfzc(work), nv)
    P = len(doub)
    with mp.get_context("spawn").Pool(cores) as pool:      # spawn: never fork a CUDA-initialised process
        res = pool.map(_det_worker, [(w0.nbf, no, par_nfzc(work), budget_s / 2, k) for k in range(cores)])
    ndet = sum(r[0] for r in res)
    t_det = max(r[1] for r in res) / ndet                 # aggregate seconds per determinant on all cores
    n_overlaps = 1 + 6 + 6 * natom + 36 * natom
    # the reference recomputes compute_all_dets for 9 overlaps per (alpha, beta) element (aats.py:714-1008)
    n_overlaps_ref = 9 * 9 * natom
    dets_per_overlap = 1 + 2 * len(sing) + 2 * P + len(sing) ** 2 + 2 * P * len(sing) + P * P
    total = npts * t_solve + n_overlaps_ref * dets_per_overlap * t_det
    desc = ("oracle (numpy port of ci_wfn.py:420-574 + aats.py:581-618) on %d host cores: %d CISD solves timed "
            "(%.3f s each, x%d points) + %d substituted %dx%d determinants timed in %d concurrent processes "
            "(%.3f us each aggregate, x%.3g dets x %d overlap evaluations as the reference recomputes them per "
            "element; %d distinct overlaps); contraction of the 8-index tensors not included (lower bound)"
            % (cores, nsolve, t_solve, npts, ndet, no, no, cores, t_det * 1e6, dets_per_overlap, n_overlaps_ref,
               n_overlaps))
    return total, desc


def par_nfzc(work):
    return work["w0"].H.basis_set.n_frozen_core()


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


def flush_l2(buf):
    buf.zero_()       # 512 MB write, larger than the 126 MB L2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="h2o2", choices=sorted(WORKLOADS))
    ap.add_argument("--profile-step", action="store_true",
                    help="run the warm-up and ONE device-resident step only (target for ncu launch lists)")
    ap.add_argument("--aat-graph", type=int, default=None, choices=[0, 1],
                    help="replay the AAT overlap stacks from CUDA graphs (default: apyib_b200.config.AAT_USE_GRAPH)")
    ap.add_argument("--aat-algorithm", default="lu", choices=["lu", "lemma", "factorized"],
                    help="substituted determinants by sub-warp LU (north star) or by the determinant lemma")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": wl["label"], "nbf": wl["nbf"], "ndocc": wl["ndocc"], "natom": wl["natom"],
              "fd_points": 6 * wl["natom"] + 7, "h_R": H_R, "h_B": H_B, "method": "CISD",
              "l2": "512 MB buffer written between timed steps (flush)", "sharding": "fd-points, then tensor rows alpha"}

    if args.impl == "reference":
        if rank != 0:
            return
        work = prepare(wl)
        vals = []
        for _ in range(max(1, min(args.steps, 2))):
            v, desc = cpu_sample(work, budget_s=30.0)
            vals.append(v)
        v = float(np.median(vals))
        cores = os.cpu_count()
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
                          "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64/c128",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import apyib_b200
    from apyib_b200 import _lib, device as dev
    apyib_b200.config.VERBOSE = False
    if args.workload == "methyloxirane" and args.aat_algorithm == "lu":
        args.aat_algorithm = "factorized"      # 2.5e10 16x16 LUs per overlap: only the closed forms are feasible
    apyib_b200.config.AAT_ALGORITHM = args.aat_algorithm
    config["aat_algorithm"] = args.aat_algorithm
    if args.aat_graph is not None:
        apyib_b200.config.AAT_USE_GRAPH = bool(args.aat_graph)
    use_graph = bool(apyib_b200.config.AAT_USE_GRAPH)
    config["aat_graph"] = use_graph
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    work = prepare(wl)
    l2buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    per_step = []

    def timed_steps(k, device_resident):
        apyib_b200.config.RETURN_DEVICE = device_resident
        tot = 0.0
        res = None
        per_step.clear()
        for _ in range(k):
            if not device_resident:
                drop_device_caches(work)
            flush_l2(l2buf)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            res = gpu_step(work, rank, world, dist)
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            per_step.append(max(e0.elapsed_time(e1) * 1e-3, wall))   # host-orchestrated step: never below wall clock
            tot += per_step[-1]
        return tot, res

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
    sampler.start()
    # with graph replay of the AAT stacks the individual launches carry no events: the step is timed as it
    # ships (graphs on) and the dominant kernel in ONE extra instrumented eager step right after it
    apyib_b200.config.TIMING = None if use_graph else {}
    # kernels timed live (CUDA events on the launching stream): the LU kernel, or for the closed-form
    # algorithms the lemma kernel and the TMA-fed contractions (ladder); launches replayed from a CUDA
    # graph cannot carry events, so for those only the eager first iteration of every solve is timed
    apyib_b200.config.TIMING_ONLY = ("det_matvec", "det_pairs") if args.aat_algorithm == "lu" else ("lemma_matvec", "contract_tma[")
    _lib.LAUNCHES[0] = 0
    t_dev, I_dev = timed_steps(args.steps, True)
    step_times = [round(x, 6) for x in per_step]
    launches = _lib.LAUNCHES[0] // max(args.steps, 1)
    timing = apyib_b200.config.TIMING
    timing_steps = args.steps
    if use_graph:
        apyib_b200.config.AAT_USE_GRAPH, apyib_b200.config.TIMING = False, {}
        timed_steps(1, True)
        timing, timing_steps = apyib_b200.config.TIMING, 1
        apyib_b200.config.AAT_USE_GRAPH = True
    apyib_b200.config.TIMING = None
    dev.COUNTERS["h2d_bytes"] = dev.COUNTERS["d2h_bytes"] = 0
    t_e2e, I_e2e = timed_steps(args.steps, False)
    h2d, d2h = dev.COUNTERS["h2d_bytes"] // args.steps, dev.COUNTERS["d2h_bytes"] // args.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if dist is not None:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    assert np.abs(I_dev - I_e2e).max() < 1e-8 * max(1.0, np.abs(I_e2e).max()), "device-resident and e2e legs disagree"

    # ---- roofline of the dominant kernel: the fused LU determinant kernel, largest shape ----
    peaks, peak_src = load_peaks()
    fp64_peak = FP64_PEAK_TFLOPS_FALLBACK
    try:
        import ctypes as C
        fl, ms = C.c_double(), C.c_float()
        _lib.check(_lib.lib.apyib_peak_fp64(1, 4000, C.byref(fl), C.byref(ms)))
        fp64_peak = fl.value / 1e12
    except Exception:
        pass
    roof = None
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except Exception:
        pass
    if timing:
        import re
        tot_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in timing.items()}
        if args.aat_algorithm == "lu":
            name = max(tot_ms, key=tot_ms.get)
        else:   # graph-replayed launches are not timed: rank by the duration of ONE launch instead of the sum
            name = max(tot_ms, key=lambda k: tot_ms[k] / len(timing[k]))
        evs = timing[name]
        avg_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        share = (sum(a.elapsed_time(b) for a, b in evs) * 1e-3 / timing_steps) / (t_dev / args.steps)
        n = wl["ndocc"]
        extra = {}
        if name.startswith("det_pairs"):
            # prefix-shared LU (csrc/dets_pairs.cu): one pivoted LU of the n-k unsubstituted columns per (row list,
            # group), one Schur k-vector per candidate column, k x k determinants per list
            k_sub = int(name.split("k=")[1].split(",")[0])
            nrow, ncol = (int(x) for x in name.split(",")[2].rstrip("]").split("x"))
            n_stack = int(name.split("nS=")[1].rstrip("]")) if "nS=" in name else 1
            ndet = nrow * ncol * n_stack
            nc = wl["nbf"] - n
            gl = nc if k_sub == 1 else nc * (nc - 1) // 2
            npre = n - k_sub
            cmac = (sum((n - 1 - j) * (npre - j) for j in range(npre))          # prefix LU: L column + trailing update
                    + k_sub * npre * (npre - 1) // 2 + k_sub * npre               # X = -L21 L11^-1
                    + nc * k_sub * npre                                          # candidate columns
                    + (gl * 4 + 2 * nc if k_sub == 2 else 2 * gl))               # k x k determinants + table x vector
            exec_flops = n_stack * nrow * (ncol // gl) * cmac * 8.0
            flops = ndet * (8.0 / 3.0) * n ** 3
            kern = "det_pairs_kernel<N=%d,K=%d> (prefix-shared LU + table x vector, %d overlaps per launch) " % (n, k_sub, n_stack)
            extra = {"determinants_per_s": ndet / (avg_ms * 1e-3), "executed_tflops": exec_flops / (avg_ms * 1e-3) / 1e12,
                     "executed_frac": exec_flops / (avg_ms * 1e-3) / 1e12 / fp64_peak,
                     "executed_complex_macs_per_determinant": cmac / gl}
            note = ("FP64 compute roofline (LU on the FP64 FMA pipe); achieved = ALGORITHMIC (8/3)n^3 flops per determinant "
                    "(SURVEY 8d U3) / measured time, so frac > 1 measures the algorithmic saving of sharing the prefix "
                    "factorisation inside a group (%.1f complex MACs per determinant instead of n^3/3 = %.0f); executed_frac "
                    "is the FP64-pipe utilisation on the flops really executed; peak = own DMMA/DFMA microbenchmark "
                    "measured in this run (MEASURED_PEAKS.json holds bf16/HBM only, %s)" % (cmac / gl, n ** 3 / 3.0, peak_src))
            tkey = "det_pairs_kernel"
        elif name.startswith("det_matvec"):
            nrow, ncol = (int(x) for x in name.split(",")[1].rstrip("]").split("x"))
            ndet = nrow * ncol * (int(name.split("nS=")[1].rstrip("]")) if "nS=" in name else 1)
            kern = "det_tpm_kernel<N=%d,B=3> (fused LU + table x vector) " % n if n <= 12 else "det_kernel<N=%d,fused> " % n
            # (8/3) n^3 real flop per substituted n x n complex LU (algorithmic count)
            flops = ndet * (8.0 / 3.0) * n ** 3
            extra = {"determinants_per_s": ndet / (avg_ms * 1e-3)}
            note = ("FP64 compute roofline (LU on the FP64 FMA pipe); achieved = ALGORITHMIC (8/3)n^3 flops per "
                    "determinant / measured time; peak = own DMMA/DFMA microbenchmark measured in this run "
                    "(MEASURED_PEAKS.json holds bf16/HBM only, %s)" % peak_src)
            if apyib_b200.config.LU_REUSE and 2 <= n <= 12 and nrow == ncol:
                # factorisation reuse: determinants whose column list shares its leading panels with the previous
                # one only redo the last panel (left-looking update + 3 pivots); count what was really executed
                from apyib_b200.aats import _Tables
                Tb = _Tables.get(n, wl["nfzc"], wl["nbf"] - n)
                cs = Tb.LS[2][0].cpu().numpy()
                nchunk = int(_lib.lib.apyib_det_matvec_work_len(nrow, ncol, 1, n)) // nrow
                clen = -(-ncol // nchunk)
                nl = ((n + 2) // 3 - 1) * 3
                same = np.concatenate([[False], (cs[1:, :nl] == cs[:-1, :nl]).all(axis=1)])
                same[np.arange(0, ncol, clen)] = False
                reuse_frac = float(same.mean())
                last = sum(n - 1 - k for k in range(nl)) * (n - nl) + sum((n - 1 - j) * (n - 1 - j) for j in range(nl, n))
                exec_flops = ndet * ((1 - reuse_frac) * (8.0 / 3.0) * n ** 3 + reuse_frac * 8.0 * last)
                extra.update({"reuse_fraction": reuse_frac, "executed_tflops": exec_flops / (avg_ms * 1e-3) / 1e12,
                              "executed_frac": exec_flops / (avg_ms * 1e-3) / 1e12 / fp64_peak})
                note += ("; with factorisation reuse %.0f %% of the determinants only redo their last panel, so the "
                         "flops actually executed are lower: executed_frac is the FP64-pipe utilisation, frac the "
                         "algorithmic rate (profiles/: plain LU without reuse = 38 %% of peak, every flop executed)"
                         % (100 * reuse_frac))
            tkey = "det_tpm_kernel"
        elif name.startswith("lemma_matvec"):
            dims = name.split(",")[1]
            nrow, ncol = (int(x) for x in dims.split("x"))
            ndet = nrow * ncol * int(name.split("nS=")[1].rstrip("]"))
            kern = "lemma_kernel<fused> "
            flops = ndet * (8.0 / 3.0) * n ** 3
            extra = {"determinants_per_s": ndet / (avg_ms * 1e-3)}
            note = ("ALGORITHMIC flops of the reference's formulation ((8/3)n^3 per n x n LU, SURVEY 8d U3) / measured "
                    "time; the lemma kernel executes far fewer flops than it is credited with, so frac is a speed-up "
                    "measure, not a pipe utilisation; peak = own FP64 microbenchmark in this run (%s)" % peak_src)
            tkey = "lemma_kernel"
        else:
            m = re.match(r"contract(_tma)?\[(f64|c128) (\d+)x(\d+)x(\d+) b(\d+)\]", name)
            M_, N_, K_, nb_ = (int(m.group(i)) for i in (3, 4, 5, 6))
            flops = (8.0 if m.group(2) == "c128" else 2.0) * M_ * N_ * K_ * nb_
            kern = "contract_tma_kernel (DMMA, TMA-fed) " if m.group(1) else "contract_kernel (DMMA) "
            note = ("FP64 tensor (DMMA) roofline, dense flop count of the reference's einsum (8 MNK complex / 2 MNK real "
                    "per point); only the eager first iteration of each solve carries events (the rest replays from a "
                    "CUDA graph), so share_of_step counts those launches only; peak = own DMMA microbenchmark in this "
                    "run (%s)" % peak_src)
            tkey = "contract_tma_kernel" if m.group(1) else "contract_kernel"
        ach = flops / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "traffic": traffic_tab.get(tkey, {}).get("dram_bytes_per_launch"), "kernel": kern + name,
                "launches_timed": len(evs), "avg_ms": avg_ms, "share_of_step": share, "note": note}
        roof.update(extra)
        if "nS=" in name and not name.endswith("nS=1]") and roof["traffic"] is not None:
            # the ncu capture is of a single-overlap launch; a stack launch adds one S and one Y per further overlap
            roof["traffic_single_overlap_launch"] = roof["traffic"]
            roof["traffic"] = None
        if use_graph:
            roof["note"] += ("; the timed steps replay the AAT stacks from CUDA graphs (no per-launch events), so the "
                             "kernel was timed in one extra eager step of the same workload right after them; "
                             "share_of_step = its launches x avg_ms over the graph-replayed step time")
        if tkey in traffic_tab:
            roof["traffic_source"] = traffic_tab[tkey].get("source")
    # the algorithmic fast path (determinant lemma) on the same workload, as an extra leg
    alt = None
    if args.aat_algorithm == "lu" and world == 1:
        apyib_b200.config.AAT_ALGORITHM = "lemma"
        timed_steps(2 if use_graph else 1, True)          # (graphs are captured at the second sight of a shape)
        ta, Ia = timed_steps(args.steps, True)
        ta_med = float(np.median(per_step))
        te, Ie = timed_steps(args.steps, False)
        te_med = float(np.median(per_step))
        apyib_b200.config.AAT_ALGORITHM = "lu"
        # informational leg: median of the steps (a first-use module load inside one step would dominate a mean of 3)
        alt = {"aat_algorithm": "lemma", "value": ta_med, "e2e": te_med, "unit": UNIT, "statistic": "median of steps",
               "max_abs_diff_vs_lu": float(np.abs(Ia - I_dev).max())}
    if args.workload == "methyloxirane":
        cpu_v, cpu_desc = None, "not runnable: the reference materialises an 8 TB determinant tensor (aats.py:575)"
    else:
        cpu_v, cpu_desc = cpu_sample(work, budget_s=20.0)
    line = {"metric": METRIC, "value": t_dev / args.steps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64/c128", "data": "synthetic", "config": config,
            "clocks": sampler.result(),
            "e2e": {"value": t_e2e / args.steps, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "roofline": roof,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": cpu_desc},
            "alt": alt, "step_times_s": step_times, "aat_checksum": float(np.abs(I_dev).sum())}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
