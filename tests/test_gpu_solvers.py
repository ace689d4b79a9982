"""GPU parity of the drop-in solver classes against the oracle (numpy restatement of the
reference) on the same seeded synthetic inputs.  Tolerances are the north star's:
energies 1e-10 Eh, amplitudes 1e-9 max-abs."""
import os

import numpy as np
import pytest

from oracle import apyib_oracle as orc

pytestmark = pytest.mark.gpu

E_TOL, T_TOL = 1e-10, 1e-9


def par(method, fc=False, maxit=120, diis=True, conv=1e-12):
    return {"method": method, "freeze_core": fc, "DIIS": diis, "max_iterations": maxit,
            "e_convergence": conv, "d_convergence": conv}


CASES = [  # nbf, ndocc, nfzc, complex
    (7, 5, 0, False), (7, 5, 1, True), (6, 2, 0, True), (10, 3, 1, False),
]


@pytest.mark.parametrize("nbf,no,nf,cplx", CASES)
def test_mo_transform_and_fock(nbf, no, nf, cplx):
    from apyib_b200 import utils
    w = orc.rotated_wfn(nbf, no, 21, cplx, nf)
    p = par("CISD", nf > 0)
    C_list, _ = orc.get_slices(p, w)
    C2, I2 = utils.get_slices(p, w)
    assert C2 == C_list
    F, Efc = utils.compute_F_MO(p, w, C_list)
    Fo, Efo = orc.compute_F_MO(p, w, C_list)
    assert np.abs(F - Fo).max() < 1e-12 and abs(Efc - Efo) < 1e-12
    assert F.dtype == Fo.dtype
    G = utils.compute_ERI_MO(p, w, C_list)
    Go = orc.compute_ERI_MO(p, w, C_list)
    assert np.abs(G - Go).max() < 1e-13 and G.dtype == Go.dtype


@pytest.mark.parametrize("nbf,no,nf,cplx", CASES)
@pytest.mark.parametrize("method", ["MP2", "MP2_SO"])
def test_mp2(nbf, no, nf, cplx, method):
    import apyib_b200
    w = orc.rotated_wfn(nbf, no, 31, cplx, nf)
    p = par(method, nf > 0)
    m = apyib_b200.mp2_wfn(p, w)
    E, t2 = m.solve_MP2() if method == "MP2" else m.solve_MP2_SO()
    Eo, t2o = orc.solve_MP2(p, w) if method == "MP2" else orc.solve_MP2_SO(p, w)
    assert abs(E - Eo) < E_TOL and np.abs(t2 - t2o).max() < T_TOL
    assert t2.shape == t2o.shape and t2.dtype == t2o.dtype
    assert np.array_equal(m.D_ijab, orc._denoms(w.eps[m.C_list[1]], w.eps[m.C_list[2]])[1])


@pytest.mark.parametrize("nbf,no,nf,cplx", CASES)
@pytest.mark.parametrize("method", ["CID", "CID_SO", "CISD", "CISD_SO"])
@pytest.mark.parametrize("diis", [True, False])
def test_ci_converged(nbf, no, nf, cplx, method, diis):
    import apyib_b200
    w = orc.rotated_wfn(nbf, no, 41, cplx, nf)
    p = par(method, nf > 0, diis=diis)
    ci = apyib_b200.ci_wfn(p, w)
    got = getattr(ci, "solve_" + method)()
    want = getattr(orc, "solve_" + method)(p, w, True)
    assert ci.iterations == want[-1], "iteration count differs from the reference semantics"
    assert abs(got[0] - want[0]) < E_TOL
    for a, b in zip(got[1:], want[1:-1]):
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.abs(a - b).max() < T_TOL
    assert type(got[0]) == type(want[0]) or np.asarray(got[0]).dtype == np.asarray(want[0]).dtype


@pytest.mark.parametrize("method", ["CID", "CISD_SO", "CISD"])
@pytest.mark.parametrize("graph", [True, False])
def test_ci_fixed_iterations_no_early_exit(method, graph):
    """max_iterations=5 with zero thresholds: the whole iteration path (incl. DIIS) must agree."""
    import apyib_b200
    apyib_b200.config.USE_CUDA_GRAPH = graph
    try:
        w = orc.rotated_wfn(8, 3, 51, True, 0)
        p = par(method, False, maxit=5, conv=0.0)
        ci = apyib_b200.ci_wfn(p, w)
        got = getattr(ci, "solve_" + method)()
        want = getattr(orc, "solve_" + method)(p, w)
        assert abs(got[0] - want[0]) < E_TOL
        for a, b in zip(got[1:], want[1:]):
            assert np.abs(a - b).max() < T_TOL
    finally:
        apyib_b200.config.USE_CUDA_GRAPH = True


def test_ci_more_than_eight_diis_vectors():
    """history truncation (utils.py:109-112) is exercised when > 8 iterations are needed"""
    import apyib_b200
    w = orc.rotated_wfn(8, 3, 61, False, 0, scale=0.03)
    p = par("CISD", False, maxit=14, conv=0.0)
    got = apyib_b200.ci_wfn(p, w).solve_CISD()
    want = orc.solve_CISD(p, w)
    assert abs(got[0] - want[0]) < E_TOL
    assert np.abs(got[2] - want[2]).max() < T_TOL


def test_ci_attributes_match_reference_layout():
    import apyib_b200
    w = orc.rotated_wfn(7, 3, 71, True, 1)
    p = par("CISD", True)
    ci = apyib_b200.ci_wfn(p, w)
    o = orc._CI(p, w)
    assert np.abs(ci.F_MO - o.F_MO).max() < 1e-12
    assert np.abs(ci.ERI_MO - o.ERI_MO).max() < 1e-13
    assert np.array_equal(ci.D_ia, o.D_ia) and np.array_equal(ci.D_ijab, o.D_ijab)


@pytest.mark.parametrize("method", ["CID", "CID_SO", "CISD", "CISD_SO"])
def test_batched_solve_equals_single_solves(method):
    """solve_many (shared launches, per-point freeze at convergence) == one solve per point"""
    import apyib_b200
    from apyib_b200.ci_wfn import solve_many
    # different points converge after different iteration counts (scales differ)
    ws = [orc.rotated_wfn(7, 3, 300 + k, False, 0, scale=sc) for k, sc in enumerate((0.005, 0.02, 0.01))]
    ws += [orc.rotated_wfn(7, 3, 310 + k, True, 0, scale=sc) for k, sc in enumerate((0.02, 0.008))]
    p = par(method)
    got = solve_many(method, p, ws)
    from apyib_b200.ci_wfn import solve_batch
    cis = [apyib_b200.ci_wfn(p, w) for w in ws]
    shared = {}
    for grp in ([0, 1, 2], [3, 4]):          # same MO integrals -> shared launches must change nothing at all
        res, its = solve_batch(method, p, [cis[k].point() for k in grp])
        for k, r, it in zip(grp, res, its):
            shared[k] = (r, it)
    for k, (w, g) in enumerate(zip(ws, got)):
        want = getattr(orc, "solve_" + method)(p, w)
        assert abs(g[0] - want[0]) < E_TOL
        for a, b in zip(g[1:], want[1:]):
            assert a.dtype == b.dtype and np.abs(a - b).max() < T_TOL
        single = getattr(cis[k], "solve_" + method)()
        for a, b in zip(shared[k][0], single):
            assert np.array_equal(np.asarray(a), np.asarray(b)), "batched and single iterations must be bit-identical"
        assert shared[k][1] == cis[k].iterations
        # solve_many also builds the MO integrals of a group with shared launches (different tile schedule):
        for a, b in zip(g, single):
            assert np.abs(np.asarray(a) - np.asarray(b)).max() < 1e-12


def test_stacked_mo_integrals_match_per_point_constructors():
    """ci_wfn.many (AO->MO transform + Fock build of a stack of points by shared launches, utils.py:217-279)
    vs one constructor per point and vs the oracle; frozen core, real and complex points mixed"""
    import apyib_b200
    ws = [orc.rotated_wfn(8, 3, 400 + k, k >= 3, 1) for k in range(5)]
    p = par("CISD", True)
    many = apyib_b200.ci_wfn.many(p, ws)
    assert many[0]._ERI_dev.data_ptr() + many[0]._ERI_dev.numel() * 8 == many[1]._ERI_dev.data_ptr()   # one stack
    for w, c in zip(ws, many):
        one = apyib_b200.ci_wfn(p, w)
        o = orc._CI(p, w)
        assert c.F_MO.dtype == one.F_MO.dtype == o.F_MO.dtype and c.ERI_MO.shape == o.ERI_MO.shape
        assert np.abs(c.F_MO - o.F_MO).max() < 1e-12 and np.abs(c.ERI_MO - o.ERI_MO).max() < 1e-13
        assert np.abs(c.F_MO - one.F_MO).max() < 1e-13 and abs(c.E_fc - o.E_fc) < 1e-11 and abs(one.E_fc - o.E_fc) < 1e-11
        assert np.array_equal(c.D_ijab, one.D_ijab)


@pytest.mark.parametrize("nbf,no,nf,cplx,seed", [(7, 3, 0, False, 901), (6, 2, 0, True, 902), (8, 3, 1, False, 903)])
def test_perturbed_cisd_amplitudes(nbf, no, nf, cplx, seed):
    """a21 (analytic_aats.py:780-885 / 1032-1137): GPU linear-response iterations vs the oracle and vs
    finite differences of the unmodified reference solver (tests/golden/synthetic_linear_response.npz)."""
    import os
    import apyib_b200
    from apyib_b200.ci_wfn import solve_perturbed_CISD
    from golden.make_golden import perturbation
    w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
    p = par("CISD", nf > 0, maxit=300, conv=1e-14)
    dF, dG = perturbation(nbf - nf, cplx, seed + 50)
    ci = apyib_b200.ci_wfn(p, w)
    E0, t1, t2 = ci.solve_CISD()
    for guess in (0.0, 0.3):
        dE, dt1, dt2 = solve_perturbed_CISD(p, ci, t1, t2, E0, dF, dG, dE_guess=guess)
        oE, o1, o2, its = orc.solve_perturbed_CISD(p, w, t1, t2, E0, dF, dG, guess, True)
        assert ci.iterations == its
        assert abs(dE - oE) < E_TOL and np.abs(dt1 - o1).max() < T_TOL and np.abs(dt2 - o2).max() < T_TOL
        assert dt1.dtype == o1.dtype and dt2.shape == o2.shape
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synthetic_linear_response.npz"))
    name = {901: "r7", 902: "c6", 903: "r8fc"}[seed]
    assert abs(dE - G[name + "/dE"]) < 1e-8 and np.abs(dt2 - G[name + "/dt2"]).max() < 1e-8


@pytest.mark.parametrize("nbf,no,nf,cplx,seed", [(7, 3, 0, False, 901), (6, 2, 0, True, 902), (8, 3, 1, False, 903)])
def test_perturbed_cid_amplitudes(nbf, no, nf, cplx, seed):
    """a21, CID variant (analytic_aats.py:1577-1649 / 1778-1850): GPU linear-response iterations vs the oracle
    and vs finite differences of the unmodified reference solve_CID (synthetic_linear_response_cid.npz)."""
    import os
    import apyib_b200
    from apyib_b200.ci_wfn import solve_perturbed_CID
    from golden.make_golden import perturbation
    w = orc.rotated_wfn(nbf, no, seed, cplx, nf)
    p = par("CID", nf > 0, maxit=300, conv=1e-14)
    dF, dG = perturbation(nbf - nf, cplx, seed + 50)
    ci = apyib_b200.ci_wfn(p, w)
    E0, t2 = ci.solve_CID()
    for guess in (0.0, 0.3):
        dE, dt2 = solve_perturbed_CID(p, ci, t2, E0, dF, dG, dE_guess=guess)
        oE, o2, its = orc.solve_perturbed_CID(p, w, t2, E0, dF, dG, guess, True)
        assert ci.iterations == its
        assert abs(dE - oE) < E_TOL and np.abs(dt2 - o2).max() < T_TOL
        assert dt2.dtype == o2.dtype and dt2.shape == o2.shape
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synthetic_linear_response_cid.npz"))
    name = {901: "r7", 902: "c6", 903: "r8fc"}[seed]
    assert abs(dE - G[name + "/dE"]) < 1e-8 and np.abs(dt2 - G[name + "/dt2"]).max() < 1e-8


@pytest.mark.parametrize("field", [False, True])
def test_scf_with_device_jk_equals_host_scf(field):
    """config.SCF_DEVICE_JK: (2J - K)[D] of every SCF iteration on the device == the host GEMV"""
    import apyib_b200
    from apyib_b200 import hostchem as hc
    prov = hc.SyntheticProvider(12, 4, 2, seed=21)
    p = {"geom": prov.geometry_string(), "basis": "synthetic", "method": "CISD", "freeze_core": False,
         "F_el": [0.0] * 3, "F_mag": [0.0, 1e-3 if field else 0.0, 0.0], "provider": prov, "DIIS": True,
         "max_iterations": 100, "e_convergence": 1e-13, "d_convergence": 1e-13}
    cfg = apyib_b200.config
    old = cfg.SCF_DEVICE_JK
    res = []
    try:
        for flag in (False, True):
            cfg.SCF_DEVICE_JK = flag
            w = hc.hf_wfn(hc.Hamiltonian(p))
            E, C = w.solve_SCF(p)
            res.append((E, np.abs(C), w.eps))
    finally:
        cfg.SCF_DEVICE_JK = old
    assert abs(res[0][0] - res[1][0]) < 1e-11 and np.abs(res[0][2] - res[1][2]).max() < 1e-10
    assert np.abs(res[0][1] - res[1][1]).max() < 1e-8         # |C|: eigenvector phases are arbitrary
    assert np.iscomplexobj(res[1][0]) == np.iscomplexobj(res[0][0])


@pytest.mark.parametrize("cplx", [False, True])
def test_solve_general_diis_drop_in(cplx):
    """utils.solve_general_DIIS with the reference's calling convention (utils.py:104-140), incl. the > 7-vector
    truncation and the iteration == 1 seeding, against the oracle"""
    from apyib_b200.utils import solve_general_DIIS
    rng = np.random.default_rng(77)
    n = 500
    rnd = lambda *s: rng.standard_normal(s) + (1j * rng.standard_normal(s) if cplx else 0)
    e_g, t_g = rnd(n, 1), rnd(n, 1)
    e_o, t_o = e_g.copy(), t_g.copy()
    for it in range(1, 12):
        r, t = (e_g[:, 0], t_g[:, 0]) if it == 1 else (rnd(n), rnd(n))
        tn_g, e_g, t_g = solve_general_DIIS({}, r, t, e_g, t_g, it)
        tn_o, e_o, t_o = orc.solve_general_DIIS(r, t, e_o, t_o, it)
        assert e_g.shape == e_o.shape and e_g.shape[1] <= 8 and np.array_equal(e_g, e_o) and np.array_equal(t_g, t_o)
        assert tn_g.dtype == tn_o.dtype and np.abs(tn_g - tn_o).max() < 1e-9 * max(1.0, np.abs(tn_o).max())


def test_concurrent_identical_batches_equal_sequential():
    """ci_wfn.solve_batches: two batches of IDENTICAL shape and dtype solved concurrently on two streams (what the
    chunked finite-difference driver does at cc-pVDZ sizes) must equal the same batches solved one after the other --
    shapes chosen so that the T1 <- T2 couplings take the split-K path, whose scratch buffers must not be shared
    between streams"""
    import apyib_b200
    from apyib_b200.ci_wfn import ci_wfn, solve_batches
    cfg = apyib_b200.config
    p = par("CISD")
    ws = [orc.rotated_wfn(26, 6, 700 + k, False, 0) for k in range(4)]
    pts = [ci_wfn(p, w).point() for w in ws]
    old = (cfg.SOLVE_CONCURRENT, cfg.RETURN_DEVICE)
    try:
        cfg.RETURN_DEVICE = False
        res = {}
        for flag in (False, True, True, True):
            cfg.SOLVE_CONCURRENT = flag
            out = solve_batches("CISD", p, [pts[:2], pts[2:]])
            flat = [r for batch, _ in out for r in batch]
            if flag not in res:
                res[flag] = flat
            else:
                for a, b in zip(res[flag], flat):
                    assert all(np.array_equal(x, y) for x, y in zip(a, b))
        for a, b in zip(res[False], res[True]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        for w, r in zip(ws, res[True]):
            Eo, t1o, t2o = orc.solve_CISD(p, w)
            assert abs(r[0] - Eo) < E_TOL and np.abs(r[2] - t2o).max() < T_TOL
    finally:
        cfg.SOLVE_CONCURRENT, cfg.RETURN_DEVICE = old


@pytest.mark.parametrize("kind", ["H", "R"])
@pytest.mark.parametrize("cplx", [False, True])
def test_perturbed_mp2_amplitudes_and_dERI(kind, cplx):
    """a21: closed-form perturbed MP2 amplitudes (analytic_aats.py:347-352, 446-451) and the perturbed-integral builds
    (:730-733, 977-981) vs the oracle's transcription of those lines"""
    import apyib_b200
    from apyib_b200.utils import build_dERI
    nbf, no, nf = 9, 3, 1
    w = orc.rotated_wfn(nbf, no, 930 + (kind == "R"), cplx, nf)
    p = par("MP2", fc=True)
    mp = apyib_b200.mp2_wfn(p, w)
    E, t2 = mp.solve_MP2()
    rng = np.random.default_rng(931)
    n = nbf - nf
    O, V = no - nf, nbf - no
    rnd = lambda *s: rng.standard_normal(s) + (1j * rng.standard_normal(s) if cplx else 0)
    dF, dW = 0.1 * rnd(n, n), 0.05 * rnd(n, n, n, n)
    got = mp.perturbed_t2(t2, dF, dW, kind)
    want = orc.perturbed_MP2_t2(t2, dF, dW, mp.D_ijab, O, V, kind)
    assert got.shape == want.shape == (O, O, V, V) and got.dtype == want.dtype
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
    U, Wfull = 0.1 * rnd(nbf, nbf), 0.05 * rnd(nbf, nbf, nbf, nbf)
    core = 0.05 * rnd(n, n, n, n) if kind == "R" else None
    gd = build_dERI(U, Wfull, nf, kind, core)
    wd = orc.build_dERI(U, Wfull, nf, kind, core)
    assert gd.shape == (n, n, n, n) and np.abs(gd - wd).max() < 1e-12 * max(1.0, np.abs(wd).max())
