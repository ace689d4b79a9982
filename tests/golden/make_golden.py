"""Generates the committed golden fixtures.  Runs ONLY in the build container (needs
/root/reference); the GPU box and the CPU test-suite read the committed outputs.

  python tests/golden/make_golden.py

(A) reference_literals.json -- the hard-coded known-answer values of the reference's own tests
    for its (H2)_2 molecule (tests/test_008_CISD_SO.py, test_011_AAT.py, test_012_AAT_SO.py,
    test_013_AAT_parallel.py), extracted from the test sources with `ast` (numbers are data;
    no reference code is copied).  Everything else in the reference's test-suite needs p/d
    functions and therefore a real Psi4.
(B) synthetic_*.npz -- outputs of the UNMODIFIED reference classes (stub-imported through
    oracle/ref_harness.py) on the seeded synthetic inputs of oracle.apyib_oracle, for the
    solvers, compute_all_dets, compute_spatial_aats and compute_SO_aats.
"""
from __future__ import annotations

import ast
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import apyib_oracle as orc, ref_harness   # noqa: E402

REF_TESTS = "/root/reference/apyib/tests"


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


# ---------------------------------------------------------------------------------------------
def extract_literals():
    cases = []
    for fname in ("test_008_CISD_SO.py", "test_011_AAT.py", "test_012_AAT_SO.py", "test_013_AAT_parallel.py"):
        src = open(os.path.join(REF_TESTS, fname)).read()
        tree = ast.parse(src)
        for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
            seg = ast.get_source_segment(src, fn)
            if '(H2)_2' not in seg:
                continue
            case = {"file": fname, "test": fn.name, "line": fn.lineno, "arrays": {}, "skipped_in_reference":
                    any("skip" in ast.unparse(d) for d in fn.decorator_list)}
            env = {"np": np}
            for node in ast.walk(fn):
                if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
                    name = node.targets[0].id
                    if name == "parameters" and isinstance(node.value, ast.Dict):
                        for k, v in zip(node.value.keys, node.value.values):
                            key = ast.literal_eval(k)
                            if key != "geom":
                                case.setdefault("parameters", {})[key] = ast.literal_eval(v)
                    elif name.endswith("_ref") or name in ("psi4_CISD",):
                        try:
                            val = eval(compile(ast.Expression(node.value), "<lit>", "eval"), env)
                        except Exception:
                            continue
                        env[name] = val
                        case["arrays"][name] = np.asarray(val).tolist()
                if isinstance(node, ast.Call):
                    f = node.func
                    attr = f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", "")
                    if attr in ("compute_parallel_aats", "compute_AAT"):
                        nums = [ast.literal_eval(a) for a in node.args if isinstance(a, ast.Constant)]
                        if len(nums) >= 2:
                            case["h_R"], case["h_B"] = nums[-2], nums[-1]
                        if attr == "compute_parallel_aats":
                            case["route"] = "parallel"
                    if attr in ("compute_parallel_aats", "compute_spatial_aats", "compute_SO_aats"):
                        case.setdefault("normalization", "full")
                        for kw in node.keywords:
                            if kw.arg == "normalization":
                                case["normalization"] = ast.literal_eval(kw.value)
                        if attr != "compute_parallel_aats":
                            case["route"] = attr
            cases.append(case)
    geom = None
    src = open("/root/reference/apyib/data/molecules.py").read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "H2_2":
            geom = ast.literal_eval(node.value)
    return {"molecule": "(H2)_2", "geom": geom, "source": "apyib/data/molecules.py:46-55", "cases": cases}


# ---------------------------------------------------------------------------------------------
def par(method, fc=False, maxit=120, diis=True, conv=1e-12):
    return {"method": method, "freeze_core": fc, "DIIS": diis, "max_iterations": maxit,
            "e_convergence": conv, "d_convergence": conv}


SOLVER_CASES = [  # name, nbf, ndocc, nfzc, complex, seed
    ("r7", 7, 3, 0, False, 11), ("c7fc", 7, 3, 1, True, 12), ("c6", 6, 2, 0, True, 13), ("r8fc", 8, 3, 1, False, 14),
]
AAT_SPATIAL = [("MP2", 5, 2, 0, 101), ("CID", 6, 3, 1, 102), ("CISD", 5, 2, 0, 103), ("CISD", 6, 3, 1, 104),
               ("RHF", 5, 2, 0, 105)]
AAT_SO = [("MP2_SO", 3, 1, 0, 201), ("CISD_SO", 3, 1, 0, 202), ("CID_SO", 4, 1, 0, 203), ("RHF", 3, 1, 0, 204)]


def synthetic(ns):
    out = {}
    for name, nbf, no, nf, cplx, seed in SOLVER_CASES:
        w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
        for m in ("MP2", "MP2_SO"):
            r = ns.mp2_wfn.mp2_wfn(par(m, nf > 0), w)
            E, t2 = r.solve_MP2() if m == "MP2" else r.solve_MP2_SO()
            out["%s/%s/E" % (name, m)], out["%s/%s/t2" % (name, m)] = np.asarray(E), t2
        for m in ("CID", "CID_SO", "CISD", "CISD_SO"):
            r = ns.ci_wfn.ci_wfn(par(m, nf > 0), w)
            res = quiet(getattr(r, "solve_" + m))
            for k, v in zip(("E", "t1", "t2") if len(res) == 3 else ("E", "t2"), res):
                out["%s/%s/%s" % (name, m, k)] = np.asarray(v)
            # fixed 4 iterations, no early exit, DIIS on: pins the iteration path itself
            r = ns.ci_wfn.ci_wfn(par(m, nf > 0, maxit=4, conv=0.0), w)
            res = quiet(getattr(r, "solve_" + m))
            out["%s/%s/E_it4" % (name, m)] = np.asarray(res[0])
            out["%s/%s/t2_it4" % (name, m)] = np.asarray(res[-1])
        r = ns.ci_wfn.ci_wfn(par("CISD", nf > 0), w)
        out["%s/F_MO" % name], out["%s/ERI_MO" % name] = r.F_MO, r.ERI_MO
    np.savez_compressed(os.path.join(HERE, "synthetic_solvers.npz"), **out)

    out = {}
    for (nbf, no, nf, seed) in ((5, 2, 0, 301), (6, 3, 1, 302), (7, 3, 0, 303)):
        A = orc.synthetic_aat_inputs("CISD", nbf, no, nf, 1, seed, h=1e-2)
        R = ref_harness.make_ref_aat(ns, A)
        res = R.compute_all_dets(A.overlap_pp[1][2])
        for k, v in enumerate(res):
            out["dets_%d_%d_%d/%d" % (nbf, no, nf, k)] = np.asarray(v)
    for method, nbf, no, nf, seed in AAT_SPATIAL:
        for norm in ("full", "intermediate"):
            A = orc.synthetic_aat_inputs(method, nbf, no, nf, 1, seed, h=1e-3)
            R = ref_harness.make_ref_aat(ns, A)
            vals = [[quiet(R.compute_spatial_aats, a, b, norm) for b in range(3)] for a in range(3)]
            out["spatial/%s_%d_%d_%d_%s" % (method, nbf, no, nf, norm)] = np.array(vals)
    for method, nbf, no, nf, seed in AAT_SO:
        for norm in ("full", "intermediate"):
            vals = []
            for (a, b) in ((0, 0), (1, 2), (2, 1)):
                A = orc.synthetic_aat_inputs(method, nbf, no, nf, 1, seed, h=1e-3)   # RHF route mutates overlaps
                R = ref_harness.make_ref_aat(ns, A)
                vals.append(quiet(R.compute_SO_aats, a, b, norm))
            out["so/%s_%d_%d_%d_%s" % (method, nbf, no, nf, norm)] = np.array(vals)
    S = orc.synthetic_aat_inputs("CISD_SO", 4, 2, 0, 1, 401, h=0.2).overlap_uu
    R = ref_harness.make_ref_aat(ns, orc.synthetic_aat_inputs("CISD_SO", 4, 2, 0, 1, 401, h=0.2))
    idx = [([], []), ([0, 5], []), ([], [1, 6]), ([0, 4, 1, 5], [2, 6]), ([0, 4, 0, 5], [1, 7, 3, 7]), ([2, 6, 3, 6], [0, 4, 1, 5])]
    out["so_det/S"] = S
    out["so_det/vals"] = np.array([R.compute_SO_det(S, b, k) for b, k in idx])
    out["so_det/idx"] = np.array([json.dumps(x) for x in idx])
    np.savez_compressed(os.path.join(HERE, "synthetic_aat.npz"), **out)


PERT_CASES = [("r7", 7, 3, 0, False, 901), ("c6", 6, 2, 0, True, 902), ("r8fc", 8, 3, 1, False, 903)]


def perturbation(n, cplx, seed):
    """Seeded perturbed MO integrals (dF Hermitian, dERI with the MO-integral symmetries)."""
    rng = np.random.default_rng(seed)
    dF = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
    dF = 0.1 * (dF + dF.conj().T)
    dG = 0.01 * (rng.standard_normal((n,) * 4) + (0.1j * rng.standard_normal((n,) * 4) if cplx else 0))
    dG = dG + dG.transpose(2, 3, 0, 1)
    dG = dG + dG.transpose(1, 0, 3, 2).conj()
    return dF, dG


def linear_response(ns):
    """a21: the reference's analytic perturbed-amplitude loop (analytic_aats.py:780-885) needs Psi4
    derivative integrals and cannot run here.  The fixture instead holds CENTRAL FINITE DIFFERENCES of
    the UNMODIFIED reference solve_CISD under F_MO + lam dF, ERI_MO + lam dERI: the quantity the
    linear-response iteration converges to."""
    out = {}
    lam = 1e-4
    for name, nbf, no, nf, cplx, seed in PERT_CASES:
        w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
        p = par("CISD", nf > 0, maxit=300, conv=1e-14)
        dF, dG = perturbation(nbf - nf, cplx, seed + 50)

        def solve(l):
            r = ns.ci_wfn.ci_wfn(p, w)
            r.F_MO = r.F_MO + l * dF
            r.ERI_MO = r.ERI_MO + l * dG
            return quiet(r.solve_CISD)
        (Ep, t1p, t2p), (Em, t1m, t2m), (E0, t1, t2) = solve(lam), solve(-lam), solve(0.0)
        out[name + "/dE"] = np.asarray((Ep - Em) / (2 * lam))
        out[name + "/dt1"] = (t1p - t1m) / (2 * lam)
        out[name + "/dt2"] = (t2p - t2m) / (2 * lam)
        out[name + "/E0"], out[name + "/t1"], out[name + "/t2"] = np.asarray(E0), t1, t2
    np.savez_compressed(os.path.join(HERE, "synthetic_linear_response.npz"), **out)


def linear_response_cid(ns):
    """a21, CID variant (analytic_aats.py:1577-1649): central finite differences of the UNMODIFIED
    reference solve_CID under F_MO + lam dF, ERI_MO + lam dERI (same cases and perturbations)."""
    out = {}
    lam = 1e-4
    for name, nbf, no, nf, cplx, seed in PERT_CASES:
        w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
        p = par("CID", nf > 0, maxit=300, conv=1e-14)
        dF, dG = perturbation(nbf - nf, cplx, seed + 50)

        def solve(l):
            r = ns.ci_wfn.ci_wfn(p, w)
            r.F_MO = r.F_MO + l * dF
            r.ERI_MO = r.ERI_MO + l * dG
            return quiet(r.solve_CID)
        (Ep, t2p), (Em, t2m), (E0, t2) = solve(lam), solve(-lam), solve(0.0)
        out[name + "/dE"] = np.asarray((Ep - Em) / (2 * lam))
        out[name + "/dt2"] = (t2p - t2m) / (2 * lam)
        out[name + "/E0"], out[name + "/t2"] = np.asarray(E0), t2
    np.savez_compressed(os.path.join(HERE, "synthetic_linear_response_cid.npz"), **out)


def fd_drivers(lit):
    """a13 + SURVEY 8(f).4: the UNMODIFIED reference's finite_difference.compute_APT / compute_Hessian /
    compute_Nuclear_Gradient / compute_Magnetic_Field_Gradient (fin_diff.py:27-263, 376-510) on its (H2)_2
    molecule, STO-3G, run through oracle/mini_psi4.py (s-Gaussian integrals).  Numbers only."""
    ns = ref_harness.load(with_mini_psi4=True)
    out = {"molecule": "(H2)_2", "basis": "STO-3G", "cases": []}
    for method, h_R, h_F, h_B in (("CISD", 1e-3, 1e-4, 1e-4), ("MP2", 1e-3, 1e-4, 1e-4), ("CID", 1e-3, 1e-4, 1e-4)):
        mk = lambda: {"geom": lit["geom"], "basis": "STO-3G", "method": method, "freeze_core": False, "DIIS": True,
                      "e_convergence": 1e-13, "d_convergence": 1e-13, "max_iterations": 120,
                      "F_el": [0.0, 0.0, 0.0], "F_mag": [0.0, 0.0, 0.0]}
        p = mk()
        E_list, T_list, C, basis = quiet(ns.energy.energy, p)
        fd = ns.fin_diff.finite_difference(p, basis, C)
        case = {"method": method, "h_R": h_R, "h_F": h_F, "h_B": h_B, "E_tot": float(np.real(E_list[0] + E_list[1] + E_list[2]))}
        case["APT"] = np.asarray(quiet(fd.compute_APT, h_R, h_F)).tolist()
        if method != "CID":
            case["Hessian"] = np.asarray(quiet(fd.compute_Hessian, h_R)).tolist()
        g = quiet(fd.compute_Nuclear_Gradient, h_R)
        case["nuclear_gradient"] = np.asarray(g[0]).tolist()
        g = quiet(fd.compute_Magnetic_Field_Gradient, h_B)
        case["magnetic_gradient"] = np.asarray(g[0]).tolist()
        t2p = np.asarray(g[5][2][2])
        case["mag_pos_T2_z_abs_sum"] = float(np.abs(t2p).sum())
        out["cases"].append(case)
    json.dump(out, open(os.path.join(HERE, "reference_fd_drivers.json"), "w"), indent=1)


def vcd_handoff(lit):
    """BASELINE configs[2] end to end on the reference's (H2)_2 molecule (STO-3G, MP2 and CISD): the UNMODIFIED
    reference's compute_Hessian + compute_APT + compute_parallel_aats feed its own vcd.compute_vcd_from_input
    (vcd.py:32-136) -> frequencies, IR intensities, VCD rotational strengths.  Numbers only."""
    ns = ref_harness.load(with_mini_psi4=True)
    fd = json.load(open(os.path.join(HERE, "reference_fd_drivers.json")))
    out = {"molecule": "(H2)_2", "basis": "STO-3G", "cases": []}
    for method, h_aat in (("MP2", 1e-4), ("CISD", 1e-6)):
        c = [c for c in fd["cases"] if c["method"] == method][0]
        mk = lambda: {"geom": lit["geom"], "basis": "STO-3G", "method": method, "freeze_core": False, "DIIS": True,
                      "e_convergence": 1e-13, "d_convergence": 1e-13, "max_iterations": 120,
                      "F_el": [0.0, 0.0, 0.0], "F_mag": [0.0, 0.0, 0.0]}
        hess, apt = np.array(c["Hessian"]), np.array(c["APT"])
        aat = np.asarray(quiet(ns.parallel.compute_parallel_aats, mk(), h_aat, h_aat, "full", 2))
        with np.errstate(invalid="ignore"):
            w, D, R = quiet(ns.vcd.vcd(mk()).compute_vcd_from_input, hess, apt, aat)
        out["cases"].append({"method": method, "h_R": c["h_R"], "h_F": c["h_F"], "h_aat": h_aat, "AAT": aat.tolist(),
                             "frequencies_cm1": [None if not np.isfinite(x) else float(x) for x in w],
                             "ir_intensities_kmmol": np.asarray(D).tolist(),
                             "rotational_strengths": np.asarray(R).tolist()})
    json.dump(out, open(os.path.join(HERE, "reference_vcd.json"), "w"), indent=1)


if __name__ == "__main__":
    if "--only-vcd" in sys.argv:
        vcd_handoff(json.load(open(os.path.join(HERE, "reference_literals.json"))))
        sys.exit(0)
    if "--only-fd-drivers" in sys.argv:
        fd_drivers(json.load(open(os.path.join(HERE, "reference_literals.json"))))
        sys.exit(0)
    if "--only-cid-response" in sys.argv:
        linear_response_cid(ref_harness.load())
        sys.exit(0)
    lit = extract_literals()
    json.dump(lit, open(os.path.join(HERE, "reference_literals.json"), "w"), indent=1)
    print("literals:", [(c["test"], sorted(c["arrays"])) for c in lit["cases"]])
    ns = ref_harness.load()
    synthetic(ns)
    linear_response(ns)
    linear_response_cid(ns)
    fd_drivers(lit)
    vcd_handoff(lit)
    print("wrote", sorted(os.listdir(HERE)))
