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


# -------------------------------------------------------------------------------------------------
# BASELINE.json configs[2-3] shape: (S)-methyloxirane/cc-pVDZ, nbf = 86, ndocc = 16, 4 frozen core orbitals
# (o/v = 12/70, 16 x 16 substituted determinants, 159 390 x 159 390 doubles x doubles table per overlap).
# The reference itself cannot run this size (8 TB tensor, aats.py:575); the solvers are checked against the
# oracle directly, one AAT element against the streamed oracle (oracle/sparse_aat.py: the reference's sums
# evaluated on the support of sparse amplitudes -- the GPU path runs its full dense machinery on the same
# inputs), and a sampled block of the determinant table against numpy.linalg.det.
# -------------------------------------------------------------------------------------------------
T_NBF, T_NDOCC, T_NFZC = 86, 16, 4


@pytest.mark.parametrize("cplx", [False, True])
def test_cisd_target_shape_vs_oracle(cplx):
    """ci_wfn.py:420-574 at (o, v) = (12, 70), frozen core: unbatched AO->MO transform (nbf^5), TMA ladder tiles,
    graph-replayed iterations; energies 1e-10, amplitudes 1e-9, identical iteration counts"""
    import apyib_b200
    w = orc.rotated_wfn(T_NBF, T_NDOCC, 8600 + int(cplx), cplx, T_NFZC, scale=0.25 / T_NBF)
    p = _par("CISD", True)
    ci = apyib_b200.ci_wfn(p, w)
    E, t1, t2 = ci.solve_CISD()
    Eo, t1o, t2o, its = orc.solve_CISD(p, w, return_iters=True)
    o, v = T_NDOCC - T_NFZC, T_NBF - T_NDOCC
    assert t2.shape == (o, o, v, v) and t2.dtype == t2o.dtype and t1.shape == (o, v)
    assert ci.iterations == its
    assert abs(E - Eo) < 1e-10 and np.abs(t1 - t1o).max() < 1e-9 and np.abs(t2 - t2o).max() < 1e-9


@pytest.mark.parametrize("method,seed", [("CISD", 8611), ("CISD", 8612), ("CID", 8613)])
def test_aat_element_target_shape_vs_streamed_oracle(method, seed):
    """one (alpha, beta) element of the FD-AAT tensor at the target shape, all nine I_xy terms, closed-form
    ("factorized", what bench.py runs at this size) and determinant-lemma kernels against the streamed oracle"""
    import apyib_b200
    from apyib_b200 import aats
    from apyib_b200.aats import AAT
    from oracle import sparse_aat as sp
    cfg = apyib_b200.config
    A = sp.sparse_aat_inputs(method, T_NBF, T_NDOCC, T_NFZC, 1, seed, h=1e-4, nnz2=40, nnz1=30)
    want = {norm: sp.spatial_aat_terms_streamed(A, 1, 2, norm) for norm in ("full", "intermediate")}
    old = (cfg.AAT_ALGORITHM, cfg.AAT_USE_GRAPH)
    try:
        for algo in ("factorized", "lemma"):
            cfg.AAT_ALGORITHM = algo
            aats._block_graphs.clear()
            G = AAT.from_parts(A.method, A.nbf, A.ndocc, A.nfzc, A.nuc_pert_strength, A.mag_pert_strength,
                               **{k: getattr(A, k) for k in PARTS if hasattr(A, k)})
            G.prefetch_rows([1])
            for norm in ("full", "intermediate"):
                got = G._spatial_terms(1, 2, norm)
                k = 1 / (4 * A.nuc_pert_strength * A.mag_pert_strength)
                for name, ref in want[norm].items():
                    # 1e-8 a.u. on the AAT element = 1e-8 * 4 h_R h_B on every term (north star tolerance)
                    assert abs(k * np.imag(got[name] - ref)) < 1e-8 * max(1.0, abs(k * np.imag(ref))), (algo, norm, name, got[name], ref)
                tot = G.compute_spatial_aats(1, 2, norm)
                ref = k * np.imag(sum(want[norm].values()))
                assert abs(tot - ref) < 1e-8 * max(1.0, abs(ref)), (algo, norm, tot, ref)
    finally:
        cfg.AAT_ALGORITHM, cfg.AAT_USE_GRAPH = old
        aats._block_graphs.clear()


def test_sampled_determinant_block_target_shape():
    """300 x 300 sample of the 159 390 x 159 390 doubles x doubles table of one overlap at n = 16: the sub-warp LU
    kernel (det_kernel<16>, the n > 12 path) and the determinant-lemma kernel against numpy.linalg.det"""
    import ctypes as C
    import torch
    from apyib_b200._lib import lib, check
    from apyib_b200.aats import _Tables, _det_outer
    from apyib_b200.device import to_device, to_host, empty, ptr, stream_ptr
    rng = np.random.default_rng(8620)
    no, nf, nv = T_NDOCC, T_NFZC, T_NBF - T_NDOCC
    T = _Tables.get(no, nf, nv)
    assert T.n2 == 66 * 2415 and T.n1 == 12 * 70
    for h, tol in ((1e-4, 1e-12), (0.3, 1e-10)):               # finite-difference-like and heavily pivoting overlaps
        Sh = np.eye(T_NBF) + h * (rng.standard_normal((T_NBF, T_NBF)) + 0.1j * rng.standard_normal((T_NBF, T_NBF)))
        S = to_device(Sh, torch.complex128)
        ir, ic = np.sort(rng.choice(T.n2, 300, replace=False)), np.sort(rng.choice(T.n2, 300, replace=False))
        sub = lambda idx: T.doubles[idx].reshape(-1, 2, 2)
        want = orc._batched_sub_dets(Sh, no, sub(ir), sub(ic))
        scale = np.abs(want).max()
        rows = T.L[2][torch.from_numpy(ir).to(S.device)].contiguous()
        cols = T.L[2][torch.from_numpy(ic).to(S.device)].contiguous()
        got_lu = to_host(_det_outer(S, no, rows, cols))
        assert np.abs(got_lu - want).max() <= tol * scale
        prep = empty((1, int(lib.apyib_lemma_prep_len(T_NBF, no))), torch.complex128)
        check(lib.apyib_lemma_prepare(ptr(S), 1, T_NBF, no, ptr(prep), stream_ptr()))
        dr = T.doubles_dev[torch.from_numpy(ir).to(S.device)].contiguous()
        dc = T.doubles_dev[torch.from_numpy(ic).to(S.device)].contiguous()
        out = empty((1, 300, 300), torch.complex128)
        check(lib.apyib_lemma_outer(ptr(prep), 1, T_NBF, no, 2, ptr(dr), 300, 2, ptr(dc), 300, ptr(out), stream_ptr()))
        assert np.abs(to_host(out)[0] - want).max() <= 10 * tol * scale


@pytest.mark.parametrize("algo", ["lu", "lemma"])
def test_aat_element_h2o2_shape_vs_streamed_oracle(algo):
    """configs[1] at FULL size against an oracle (not only self-consistency): one (alpha, beta) element, all nine I_xy
    terms, with the batched-LU kernels (prefix-shared LU on the 2808 x 2808 doubles x doubles table, what bench.py
    --workload h2o2 runs) and the lemma kernel, vs the streamed oracle on sparse amplitudes"""
    import apyib_b200
    from apyib_b200 import aats
    from apyib_b200.aats import AAT
    from oracle import sparse_aat as sp
    cfg = apyib_b200.config
    A = sp.sparse_aat_inputs("CISD", NBF, NDOCC, 0, 1, 2230, h=1e-4, nnz2=40, nnz1=30)
    want = sp.spatial_aat_terms_streamed(A, 2, 1, "full")
    old = cfg.AAT_ALGORITHM
    try:
        cfg.AAT_ALGORITHM = algo
        aats._block_graphs.clear()
        G = AAT.from_parts(A.method, A.nbf, A.ndocc, A.nfzc, A.nuc_pert_strength, A.mag_pert_strength,
                           **{k: getattr(A, k) for k in PARTS if hasattr(A, k)})
        G.prefetch_rows([2])
        got = G._spatial_terms(2, 1, "full")
        k = 1 / (4 * A.nuc_pert_strength * A.mag_pert_strength)
        for name, ref in want.items():
            assert abs(k * np.imag(got[name] - ref)) < 1e-8 * max(1.0, abs(k * np.imag(ref))), (algo, name, got[name], ref)
    finally:
        cfg.AAT_ALGORITHM = old
        aats._block_graphs.clear()
