"""CPU: pins the oracle (numpy restatement) to (A) outputs of the UNMODIFIED reference on the
seeded synthetic inputs (tests/golden/synthetic_*.npz, made by make_golden.py) and (B) the
hard-coded known-answer values of the reference's own tests for its (H2)_2 molecule
(tests/golden/reference_literals.json)."""
import json
import os

import numpy as np
import pytest

from oracle import apyib_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SOLV = np.load(os.path.join(HERE, "golden", "synthetic_solvers.npz"))
AATG = np.load(os.path.join(HERE, "golden", "synthetic_aat.npz"))
LIT = json.load(open(os.path.join(HERE, "golden", "reference_literals.json")))

from golden.make_golden import SOLVER_CASES, AAT_SPATIAL, AAT_SO, par   # noqa: E402


@pytest.mark.parametrize("name,nbf,no,nf,cplx,seed", SOLVER_CASES)
def test_solvers_match_reference_outputs(name, nbf, no, nf, cplx, seed):
    w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
    ci = orc._CI(par("CISD", nf > 0), w)
    assert np.abs(ci.F_MO - SOLV[name + "/F_MO"]).max() < 1e-13
    assert np.abs(ci.ERI_MO - SOLV[name + "/ERI_MO"]).max() < 1e-14
    for m in ("MP2", "MP2_SO"):
        E, t2 = getattr(orc, "solve_" + m)(par(m, nf > 0), w)
        assert abs(E - SOLV["%s/%s/E" % (name, m)]) < 1e-14
        assert np.abs(t2 - SOLV["%s/%s/t2" % (name, m)]).max() < 1e-14
    for m in ("CID", "CID_SO", "CISD", "CISD_SO"):
        res = getattr(orc, "solve_" + m)(par(m, nf > 0), w)
        assert abs(res[0] - SOLV["%s/%s/E" % (name, m)]) < 1e-13
        assert np.abs(res[-1] - SOLV["%s/%s/t2" % (name, m)]).max() < 1e-13
        if len(res) == 3:
            assert np.abs(res[1] - SOLV["%s/%s/t1" % (name, m)]).max() < 1e-13
        res = getattr(orc, "solve_" + m)(par(m, nf > 0, maxit=4, conv=0.0), w)
        assert abs(res[0] - SOLV["%s/%s/E_it4" % (name, m)]) < 1e-13
        assert np.abs(res[-1] - SOLV["%s/%s/t2_it4" % (name, m)]).max() < 1e-13


@pytest.mark.parametrize("nbf,no,nf,seed", [(5, 2, 0, 301), (6, 3, 1, 302), (7, 3, 0, 303)])
def test_compute_all_dets_bit_exact_layout(nbf, no, nf, seed):
    A = orc.synthetic_aat_inputs("CISD", nbf, no, nf, 1, seed, h=1e-2)
    res = orc.compute_all_dets(A.overlap_pp[1][2], no, nf, nbf)
    for k, v in enumerate(res):
        g = AATG["dets_%d_%d_%d/%d" % (nbf, no, nf, k)]
        assert np.asarray(v).shape == g.shape
        assert np.abs(np.asarray(v) - g).max() < 1e-14
        assert np.array_equal(np.asarray(v) == 0, g == 0), "zero pattern (index tables) must be identical"


@pytest.mark.parametrize("method,nbf,no,nf,seed", AAT_SPATIAL)
@pytest.mark.parametrize("norm", ["full", "intermediate"])
def test_spatial_aat_matches_reference_outputs(method, nbf, no, nf, seed, norm):
    A = orc.synthetic_aat_inputs(method, nbf, no, nf, 1, seed, h=1e-3)
    got = np.array([[orc.compute_spatial_aats(A, a, b, norm) for b in range(3)] for a in range(3)])
    want = AATG["spatial/%s_%d_%d_%d_%s" % (method, nbf, no, nf, norm)]
    assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("method,nbf,no,nf,seed", AAT_SO)
@pytest.mark.parametrize("norm", ["full", "intermediate"])
def test_so_aat_matches_reference_outputs(method, nbf, no, nf, seed, norm):
    A = orc.synthetic_aat_inputs(method, nbf, no, nf, 1, seed, h=1e-3)
    got = np.array([orc.compute_SO_aats(A, a, b, norm) for (a, b) in ((0, 0), (1, 2), (2, 1))])
    want = AATG["so/%s_%d_%d_%d_%s" % (method, nbf, no, nf, norm)]
    assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())


def test_so_det_sequential_swap_semantics():
    S = AATG["so_det/S"]
    for js, want in zip(AATG["so_det/idx"], AATG["so_det/vals"]):
        bra, ket = json.loads(str(js))
        assert abs(orc.compute_SO_det(S, 4, bra, ket) - want) < 1e-14


def test_get_slices_tables():
    w = orc.synthetic_wfn(9, 4, 1, nfzc=1)
    C, I = orc.get_slices({"method": "CISD"}, w)
    assert C == [slice(0, 1), slice(1, 4), slice(4, 9), slice(1, 9)]
    assert I == [slice(0, 1), slice(0, 3), slice(3, 8), slice(0, 8)]
    C, I = orc.get_slices({"method": "CISD_SO"}, w)
    assert I == [slice(0, 2), slice(0, 6), slice(6, 16), slice(0, 16)]


# ---- (B) the reference's own known-answer values for (H2)_2 ---------------------------------
def _energy_case():
    return [c for c in LIT["cases"] if "psi4_CISD" in c["arrays"]][0]


def test_h2_2_cisd_energy_literal():
    """apyib/tests/test_008_CISD_SO.py:107-130: E_tot = -2.2165136315314133 (1e-11)"""
    from oracle import fd_pipeline as fp
    c = _energy_case()
    for method in ("CISD_SO", "CISD"):
        p = dict(c["parameters"], geom=LIT["geom"], method=method)
        E_list = fp.energy(p)[0]
        assert abs(E_list[0] + E_list[1] + E_list[2] - c["arrays"]["psi4_CISD"]) < 1e-11


AAT_CASES = [c for c in LIT["cases"] if "h_R" in c]


def _fast(c):
    return c["parameters"]["method"] in ("MP2", "CISD")


@pytest.mark.parametrize("c", [c for c in AAT_CASES if _fast(c)], ids=lambda c: c["file"][5:8] + "-" + c["test"])
def test_h2_2_aat_literals_spatial(c):
    _check_aat_case(c)


@pytest.mark.parametrize("c", [c for c in AAT_CASES if not _fast(c) and c["file"].startswith("test_013")],
                         ids=lambda c: c["file"][5:8] + "-" + c["test"])
def test_h2_2_aat_literals_spin_orbital(c):
    _check_aat_case(c)


def _check_aat_case(c):
    from oracle import fd_pipeline as fp
    p = dict(c["parameters"], geom=LIT["geom"], F_el=[0.0] * 3, F_mag=[0.0] * 3)
    I, T = fp.compute_parallel_aats(p, c["h_R"], c["h_B"], c["normalization"], terms=True)
    tol = 1e-8 if "mp2" in c["test"] and c["file"].startswith("test_013") and "SO" not in c["test"] else 1e-7
    assert np.abs(I - np.array(c["arrays"]["aat_ref"])).max() < tol
    if p["method"].endswith("_SO"):        # the SO route evaluates every term -> term-resolved check
        for k in ("00", "0D", "D0", "DD"):
            if "I_%s_ref" % k in c["arrays"]:
                assert np.abs(T[k] - np.array(c["arrays"]["I_%s_ref" % k])).max() < 1e-8
    else:
        for k in ("00", "DD"):
            if "I_%s_ref" % k in c["arrays"]:
                assert np.abs(T[k] - np.array(c["arrays"]["I_%s_ref" % k])).max() < 1e-8


# ---- a21: perturbed-amplitude (linear-response) CISD iterations ---------------------------------
from golden.make_golden import PERT_CASES, perturbation   # noqa: E402
LRG = np.load(os.path.join(HERE, "golden", "synthetic_linear_response.npz"))


@pytest.mark.parametrize("name,nbf,no,nf,cplx,seed", PERT_CASES)
def test_linear_response_matches_finite_differences_of_reference_solver(name, nbf, no, nf, cplx, seed):
    """analytic_aats.py:780-885 restated; pinned by central differences of the unmodified reference
    solve_CISD under perturbed MO integrals (the reference's own loop needs Psi4 derivative integrals)."""
    w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
    p = par("CISD", nf > 0, maxit=300, conv=1e-14)
    dF, dG = perturbation(nbf - nf, cplx, seed + 50)
    dE, dt1, dt2 = orc.solve_perturbed_CISD(p, w, LRG[name + "/t1"], LRG[name + "/t2"], LRG[name + "/E0"], dF, dG)
    assert abs(dE - LRG[name + "/dE"]) < 1e-8
    assert np.abs(dt1 - LRG[name + "/dt1"]).max() < 1e-8
    assert np.abs(dt2 - LRG[name + "/dt2"]).max() < 1e-8


LRG_CID = np.load(os.path.join(HERE, "golden", "synthetic_linear_response_cid.npz"))


@pytest.mark.parametrize("name,nbf,no,nf,cplx,seed", PERT_CASES)
def test_cid_linear_response_matches_finite_differences_of_reference_solver(name, nbf, no, nf, cplx, seed):
    """analytic_aats.py:1577-1649 restated (CID variant of a21); pinned by central differences of the
    unmodified reference solve_CID under perturbed MO integrals."""
    w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
    p = par("CID", nf > 0, maxit=300, conv=1e-14)
    dF, dG = perturbation(nbf - nf, cplx, seed + 50)
    dE, dt2 = orc.solve_perturbed_CID(p, w, LRG_CID[name + "/t2"], LRG_CID[name + "/E0"], dF, dG)
    assert abs(dE - LRG_CID[name + "/dE"]) < 1e-8
    assert np.abs(dt2 - LRG_CID[name + "/dt2"]).max() < 1e-8


# ---- a13 + 8(f).4: energy-only finite-difference drivers vs the UNMODIFIED reference on (H2)_2 ---------------
FDG = json.load(open(os.path.join(HERE, "golden", "reference_fd_drivers.json")))


@pytest.mark.parametrize("c", FDG["cases"], ids=lambda c: c["method"])
def test_h2_2_fd_drivers_match_reference_outputs(c):
    """fin_diff.py:151-263 (APT), :376-510 (gradients) restated in oracle/fd_pipeline.py; pinned by outputs of the
    unmodified reference run through oracle/mini_psi4.py (tests/golden/make_golden.py --only-fd-drivers).
    Tolerances: second differences of energies that agree to ~1e-13 (h_R h_F = 1e-7 -> 1e-6)."""
    from oracle import fd_pipeline as fp
    mk = lambda: {"geom": LIT["geom"], "basis": "STO-3G", "method": c["method"], "freeze_core": False, "DIIS": True,
                  "e_convergence": 1e-13, "d_convergence": 1e-13, "max_iterations": 120,
                  "F_el": [0.0] * 3, "F_mag": [0.0] * 3}
    p = mk()
    E_list, T0, C, basis, wfn = fp.energy(p)
    assert abs(E_list[0] + E_list[1] + E_list[2] - c["E_tot"]) < 1e-12
    apt = fp.compute_APT(mk(), c["h_R"], c["h_F"])
    assert apt.shape == (12, 3) and np.abs(apt - np.array(c["APT"])).max() < 2e-6
    g, pT, nT = fp.compute_Nuclear_Gradient(mk(), basis, C, c["h_R"])
    assert np.abs(g - np.array(c["nuclear_gradient"])).max() < 1e-9
    g, pT, nT = fp.compute_Magnetic_Field_Gradient(mk(), basis, C, c["h_B"])
    assert np.abs(g - np.array(c["magnetic_gradient"])).max() < 1e-8
    assert abs(np.abs(pT[2][2]).sum() - c["mag_pos_T2_z_abs_sum"]) < 1e-10


def test_h2_2_hessian_matches_reference_output():
    """fin_diff.py:27-147: 4 (3N)^2 = 576 energies (MP2 keeps the CPU suite short; the fixture also holds CISD)"""
    from oracle import fd_pipeline as fp
    c = [c for c in FDG["cases"] if c["method"] == "MP2"][0]
    p = {"geom": LIT["geom"], "basis": "STO-3G", "method": "MP2", "freeze_core": False, "DIIS": True,
         "e_convergence": 1e-13, "d_convergence": 1e-13, "max_iterations": 120, "F_el": [0.0] * 3, "F_mag": [0.0] * 3}
    H = fp.compute_Hessian(p, c["h_R"])
    assert H.shape == (12, 12) and np.abs(H - np.array(c["Hessian"])).max() < 2e-6


@pytest.mark.parametrize("method,nbf,no,nf", [("CISD", 7, 3, 1), ("CISD", 6, 3, 0), ("CID", 7, 3, 1)])
def test_streamed_aat_oracle_equals_dense_oracle(method, nbf, no, nf):
    """oracle/sparse_aat.py (used at sizes where the 8-index tensor cannot exist) == the dense oracle, term by term,
    on sparse and on dense amplitudes"""
    from oracle import sparse_aat as sp
    for A in (sp.sparse_aat_inputs(method, nbf, no, nf, 1, 11, h=1e-3, nnz2=6, nnz1=4),
              orc.synthetic_aat_inputs(method, nbf, no, nf, 1, 5, h=1e-3)):
        for norm in ("full", "intermediate"):
            d = orc.spatial_aat_terms(A, 1, 2, norm)
            s = sp.spatial_aat_terms_streamed(A, 1, 2, norm)
            for k in d:
                assert abs(d[k] - s[k]) < 1e-14 * max(1.0, abs(d[k])), (k, d[k], s[k])


def test_perturbed_mp2_oracle_conventions_agree():
    """oracle.perturbed_MP2_t2: the magnetic-field (analytic_aats.py:347-352) and nuclear (:446-451) transcriptions differ
    only in index conventions -- for a symmetric perturbed Fock matrix and the same <ab|ij> block they must coincide;
    build_dERI: the two kinds differ by the sign of the bra terms."""
    rng = np.random.default_rng(0)
    O, V = 2, 3
    n = O + V
    t2 = rng.standard_normal((O, O, V, V))
    dF = rng.standard_normal((n, n))
    dF = dF + dF.T
    dW = rng.standard_normal((n,) * 4)
    D = rng.standard_normal((O, O, V, V)) + 5
    a = orc.perturbed_MP2_t2(t2, dF, dW, D, O, V, "H")
    dWr = np.zeros_like(dW)
    dWr[:O, :O, O:, O:] = dW.swapaxes(0, 2).swapaxes(1, 3)[:O, :O, O:, O:]
    b = orc.perturbed_MP2_t2(t2, dF, dWr, D, O, V, "R")
    assert np.abs(a - b).max() < 1e-14
    U, W = rng.standard_normal((6, 6)), rng.standard_normal((6,) * 4)
    h, r = orc.build_dERI(U, W, 1, "H"), orc.build_dERI(U, W, 1, "R")
    ket = np.einsum("tr,pqts->pqrs", U[:, 1:], W[1:, 1:, :, 1:]) + np.einsum("ts,pqrt->pqrs", U[:, 1:], W[1:, 1:, 1:, :])
    assert h.shape == (5, 5, 5, 5) and np.abs((h + r) / 2 - ket).max() < 1e-13
