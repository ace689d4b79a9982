"""BASELINE.json configs[1] at FULL size (H2O2/6-31G shape: nbf = 22, ndocc = 9, o/v = 9/13, 9 x 9 substituted
determinants, 2808 x 2808 doubles x doubles table): direct parity where the oracle finishes in seconds (the
solvers), and size-independent properties where it cannot (one AAT element costs the reference 9 x 8.6e6
determinants): three independent evaluations of the determinant sums -- LU of every matrix, the determinant
lemma, and the closed-form factorisation -- must agree, with and without CUDA-graph replay."""
import numpy as np
import pytest

from oracle import apyib_oracle as orc

pytestmark = pytest.mark.gpu

NBF, NDOCC = 22, 9
PARTS = ("overlap_uu", "overlap_up", "overlap_un", "overlap_pu", "overlap_nu", "overlap_pp", "overlap_pn",
         "overlap_np", "overlap_nn", "unperturbed_T", "nuc_pos_T", "nuc_neg_T", "mag_pos_T", "mag_neg_T")


def _par(method, fc):
    return {"method": method, "freeze_core": fc, "DIIS": True, "max_iterations": 120, "e_convergence": 1e-12,
            "d_convergence": 1e-12}


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("nf", [0, 2])
def test_cisd_full_size_vs_oracle(cplx, nf):
    """ci_wfn.py:420-574 at the configs[1] shape, field-free (float64) and magnetic-field (complex128) points"""
    import apyib_b200
    w = orc.rotated_wfn(NBF, NDOCC, 2200 + nf, cplx, nf)
    p = _par("CISD", nf > 0)
    ci = apyib_b200.ci_wfn(p, w)
    E, t1, t2 = ci.solve_CISD()
    Eo, t1o, t2o, its = orc.solve_CISD(p, w, return_iters=True)
    assert t2.shape == (NDOCC - nf, NDOCC - nf, NBF - NDOCC, NBF - NDOCC) and t2.dtype == t2o.dtype
    assert ci.iterations == its
    assert abs(E - Eo) < 1e-10 and np.abs(t1 - t1o).max() < 1e-9 and np.abs(t2 - t2o).max() < 1e-9


def test_mp2_and_cid_full_size_vs_oracle():
    import apyib_b200
    w = orc.rotated_wfn(NBF, NDOCC, 2210, True, 0)
    p = _par("MP2", False)
    E, t2 = apyib_b200.mp2_wfn(p, w).solve_MP2()
    Eo, t2o = orc.solve_MP2(p, w)
    assert abs(E - Eo) < 1e-10 and np.abs(t2 - t2o).max() < 1e-9
    p = _par("CID", False)
    E, t2 = apyib_b200.ci_wfn(p, w).solve_CID()
    Eo, t2o = orc.solve_CID(p, w)
    assert abs(E - Eo) < 1e-10 and np.abs(t2 - t2o).max() < 1e-9


@pytest.mark.parametrize("method", ["CISD", "CID"])
def test_aat_element_full_size_three_algorithms_agree(method):
    """one (alpha, beta) element at full size: 21 overlaps x (1 + 2 ov + 2 P + (ov)^2 + 2 P ov + P^2) 9 x 9
    determinants, P = 2808 -- LU (with and without factorisation reuse, eager and graph-replayed), lemma and
    closed form agree to 1e-10 relative; 'intermediate' and 'full' normalisation both covered"""
    import apyib_b200
    from apyib_b200 import aats
    from apyib_b200.aats import AAT
    cfg = apyib_b200.config
    A = orc.synthetic_aat_inputs(method, NBF, NDOCC, 0, 1, 2222, h=1e-4)

    def element(norm):
        G = AAT.from_parts(A.method, A.nbf, A.ndocc, A.nfzc, A.nuc_pert_strength, A.mag_pert_strength,
                           **{k: getattr(A, k) for k in PARTS if hasattr(A, k)})
        G.prefetch_rows([1])
        return G.compute_spatial_aats(1, 2, norm)

    old = (cfg.AAT_ALGORITHM, cfg.LU_REUSE, cfg.AAT_USE_GRAPH)
    got = {}
    try:
        for name, algo, reuse, graph in (("lu", "lu", True, False), ("lu_plain", "lu", False, False),
                                         ("lemma", "lemma", True, False), ("factorized", "factorized", True, False),
                                         ("lu_graph", "lu", True, True)):
            cfg.AAT_ALGORITHM, cfg.LU_REUSE, cfg.AAT_USE_GRAPH = algo, reuse, graph
            aats._block_graphs.clear()
            reps = 3 if graph else 1                      # warm, capture, replay
            got[name] = [[element(norm) for norm in ("full", "intermediate")] for _ in range(reps)][-1]
    finally:
        cfg.AAT_ALGORITHM, cfg.LU_REUSE, cfg.AAT_USE_GRAPH = old
        aats._block_graphs.clear()
    ref = np.array(got["lu_plain"])
    assert np.all(np.isfinite(ref)) and np.abs(ref).max() > 0
    for name, v in got.items():
        assert np.abs(np.array(v) - ref).max() < 1e-10 * max(1.0, np.abs(ref).max()), (name, v, ref)
    assert got["lu_graph"] == got["lu"]                    # same kernels, same order: bit-identical
