"""GPU parity tests of the individual kernels (through the C-ABI) against numpy / the oracle."""
import ctypes as C

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(rng, shape, cplx):
    x = rng.standard_normal(shape)
    if cplx:
        x = x + 1j * rng.standard_normal(shape)
    return x


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("spec,shapes", [
    ("mk,kn->mn", [(70, 33), (33, 45)]),
    ("mk,kn->mn", [(200, 130), (130, 190)]),
    ("ijef,abef->ijab", [(5, 5, 7, 7), (7, 7, 7, 7)]),
    ("kbcj,ikac->ijab", [(4, 6, 6, 4), (4, 4, 6, 6)]),
    ("ls,mnls->mn", [(9, 9), (9, 9, 9, 9)]),
    ("jb,ijab->ia", [(4, 6), (4, 4, 6, 6)]),
    ("mnlg,gs->mnls", [(6, 6, 6, 6), (6, 5)]),
    ("mk,kn->mn", [(1, 1), (1, 1)]),
    ("ia,jb->iajb", [(3, 4), (2, 5)]),
    # skinny sides (16-wide tiles): T1 <-> T2 couplings, Fock-like terms, J/K builds, long-K split
    ("sabcj,sic->sijab", [(3, 9, 9, 11, 5), (3, 5, 11)]),
    ("sajbc,sijbc->sia", [(2, 30, 6, 30, 30), (2, 6, 6, 30, 30)]),
    ("ski,skjab->sijab", [(3, 5, 5), (3, 5, 5, 9, 9)]),
    ("sac,sijcb->sijab", [(2, 20, 20), (2, 4, 4, 20, 20)]),
    ("ls,mnls->mn", [(40, 40), (12, 30, 40, 40)]),
    ("mk,kn->mn", [(300, 5000), (5000, 7)]),
    ("mk,kn->mn", [(1, 9000), (9000, 200)]),
    # dot-product-like (contract_dot_kernel), with and without split-K
    ("sxia,sia->sx", [(2, 3, 5, 40), (2, 5, 40)]),
    ("xk,qk->xq", [(1, 70000), (1, 70000)]),
    ("xk,qk->xq", [(3, 900), (4, 900)]),
    ("sxklcd,qklcd->sxq", [(2, 1, 4, 4, 9, 9), (3, 4, 4, 9, 9)]),
])
def test_contract_matches_einsum(spec, shapes, cplx):
    from apyib_b200.contraction import contract
    from apyib_b200.device import to_device, to_host
    rng = np.random.default_rng(1)
    A, B = _rand(rng, shapes[0], cplx), _rand(rng, shapes[1], cplx)
    ref = np.einsum(spec, A, B)
    out0 = _rand(rng, ref.shape, cplx)
    dA, dB, dO = to_device(A), to_device(B), to_device(out0)
    alpha, beta = (0.5 - 0.25j, 1.5 + 0.5j) if cplx else (0.5, 1.5)
    contract(spec, dA, dB, dO, alpha, beta)
    want = alpha * ref + beta * out0
    assert np.abs(to_host(dO) - want).max() < 1e-12 * max(1.0, np.abs(want).max())
    # conjugated operands + beta = 0 must ignore (possibly NaN) previous contents
    dO.fill_(float("nan"))
    contract(spec, dA, dB, dO, 1.0, 0.0, conj_a=True, conj_b=True)
    want = np.einsum(spec, A.conj(), B.conj())
    assert np.abs(to_host(dO) - want).max() < 1e-12 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("cplx", [False, True])
def test_contract_strided_views(cplx):
    """operands that are slices / swapaxes views, as the reference feeds to opt_einsum"""
    from apyib_b200.contraction import contract
    from apyib_b200.device import to_device, to_host
    rng = np.random.default_rng(2)
    n, o = 9, 3
    E = _rand(rng, (n, n, n, n), cplx)
    t2 = _rand(rng, (o, o, n - o, n - o), cplx)
    W = E.swapaxes(1, 2)
    dE, dt = to_device(E), to_device(t2)
    dW = dE.swapaxes(1, 2)
    r = np.zeros_like(t2)
    dr = to_device(r)
    contract("abcd,ijcd->ijab", dW[o:, o:, o:, o:], dt, dr, 1.0, 0.0)
    contract("kbcj,ikac->ijab", dW[:o, o:, o:, :o], dt, dr, -1.0, 1.0)
    want = np.einsum("abcd,ijcd->ijab", W[o:, o:, o:, o:], t2) - np.einsum("kbcj,ikac->ijab", W[:o, o:, o:, :o], t2)
    assert np.abs(to_host(dr) - want).max() < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("spin", [0, 1])
def test_gather4_blocks(cplx, spin):
    from oracle import apyib_oracle as orc
    from apyib_b200.ci_wfn import w_block
    from apyib_b200.device import to_device, to_host
    rng = np.random.default_rng(3)
    n, o = 6, 2
    E = _rand(rng, (n, n, n, n), cplx)
    G = orc.spin_block_4(E) if spin else E
    W = G.swapaxes(1, 2)
    f = 2 if spin else 1
    O, V = f * o, f * (n - o)
    os_, vs = slice(0, O), slice(O, O + V)
    bd = {"o": (0, O), "v": (O, O + V)}
    sp = dict(i="o", j="o", k="o", a="v", b="v", c="v")
    dE = to_device(E)
    got = to_host(w_block(dE, "kbcj", "kbcj", sp, bd, spin, 1.0, -1.0))
    want = W[os_, vs, vs, os_] - W.swapaxes(2, 3)[os_, vs, vs, os_]
    assert np.array_equal(got, want)
    got = to_host(w_block(dE, "abij", "ijab", sp, bd, spin, 2.0, -1.0))
    want = 2.0 * W.swapaxes(0, 2).swapaxes(1, 3)[os_, os_, vs, vs] - W.swapaxes(2, 3).swapaxes(0, 2).swapaxes(1, 3)[os_, os_, vs, vs]
    assert np.abs(got - want).max() < 1e-15


def test_spin_block_bit_exact():
    from oracle import apyib_oracle as orc
    from apyib_b200 import utils
    rng = np.random.default_rng(4)
    F = _rand(rng, (5, 5), True)
    E = _rand(rng, (3, 4, 3, 4), True)
    assert np.array_equal(utils.compute_F_SO(None, F), orc.spin_block_2(F))
    assert np.array_equal(utils.compute_ERI_SO(None, E), orc.spin_block_4(E))
    assert np.array_equal(utils.compute_so_overlap(5, F), orc.spin_block_2(F))


@pytest.fixture(params=[0, 1], ids=["thread-per-matrix", "sub-warp"])
def det_kernel(request):
    """Both LU kernels behind apyib_det_outer / apyib_det_matvec (0 = default dispatch)."""
    from apyib_b200._lib import lib, check
    check(lib.apyib_det_set_kernel(request.param))
    yield request.param
    check(lib.apyib_det_set_kernel(0))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 17, 24, 32])
@pytest.mark.parametrize("extra", [5, 60])      # ns = n + extra: S inside / outside shared memory
def test_det_outer_vs_numpy(n, extra, det_kernel):
    from apyib_b200._lib import lib, check
    from apyib_b200.device import to_device, to_host, empty, ptr, stream_ptr
    if extra == 60 and (det_kernel == 1 or n > 12):
        pytest.skip("large-S variant only differs for the thread-per-matrix kernel")
    rng = np.random.default_rng(5 + n)
    ns = n + extra
    S = np.eye(ns) + 0.3 * (rng.standard_normal((ns, ns)) + 1j * rng.standard_normal((ns, ns)))
    nrow, ncol = 137, 29
    rows = np.array([rng.permutation(ns)[:n] for _ in range(nrow)], dtype=np.int32)
    cols = np.array([rng.permutation(ns)[:n] for _ in range(ncol)], dtype=np.int32)
    dS, dr, dc = to_device(S), torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    out = empty((nrow, ncol), torch.complex128)
    check(lib.apyib_det_outer(ptr(dS), ns, n, ptr(dr), nrow, ptr(dc), ncol, ptr(out), stream_ptr()))
    want = np.linalg.det(S[rows[:, None, :, None], cols[None, :, None, :]])
    got = to_host(out)
    assert np.abs(got - want).max() < 1e-11 * max(1.0, np.abs(want).max())


def test_det_singular_and_tiny_pivots(det_kernel):
    from apyib_b200._lib import lib, check
    from apyib_b200.device import to_device, to_host, empty, ptr, stream_ptr
    n = 6
    S = np.zeros((n, n), dtype=np.complex128)
    S[np.arange(n), (np.arange(n) + 1) % n] = 1.0          # cyclic permutation matrix, zero diagonal
    rows = np.arange(n, dtype=np.int32)[None, :]
    dS, dr = to_device(S), torch.from_numpy(rows).cuda()
    out = empty((1, 1), torch.complex128)
    check(lib.apyib_det_outer(ptr(dS), n, n, ptr(dr), 1, ptr(dr), 1, ptr(out), stream_ptr()))
    assert abs(to_host(out)[0, 0] - np.linalg.det(S)) < 1e-14
    S[:, 0] = 0.0                                          # exactly singular
    dS = to_device(S)
    check(lib.apyib_det_outer(ptr(dS), n, n, ptr(dr), 1, ptr(dr), 1, ptr(out), stream_ptr()))
    assert to_host(out)[0, 0] == 0


@pytest.mark.parametrize("n", [2, 4, 7, 9, 10, 12, 16])
def test_det_matvec_vs_numpy(n, det_kernel):
    from apyib_b200._lib import lib, check
    from apyib_b200.device import to_device, to_host, empty, ptr, stream_ptr
    rng = np.random.default_rng(50 + n)
    ns = n + 6
    S = np.eye(ns) + 0.2 * (rng.standard_normal((ns, ns)) + 1j * rng.standard_normal((ns, ns)))
    nrow, ncol, ny = 153, 211, 3
    rows = np.array([rng.permutation(ns)[:n] for _ in range(nrow)], dtype=np.int32)
    cols = np.array([rng.permutation(ns)[:n] for _ in range(ncol)], dtype=np.int32)
    Y = rng.standard_normal((ny, ncol)) + 1j * rng.standard_normal((ny, ncol))
    dS, dr, dc, dY = to_device(S), torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), to_device(Y)
    Z = empty((ny, nrow), torch.complex128)
    work = empty((int(lib.apyib_det_matvec_work_len(nrow, ncol, ny, n)),), torch.complex128)
    check(lib.apyib_det_matvec(ptr(dS), ns, n, ptr(dr), nrow, ptr(dc), ncol, ptr(dY), ny, ptr(Z), ptr(work), stream_ptr()))
    D = np.linalg.det(S[rows[:, None, :, None], cols[None, :, None, :]])
    want = Y @ D.T
    assert np.abs(to_host(Z) - want).max() < 1e-11 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("n,nv", [(4, 5), (7, 6), (9, 5), (12, 4)])
def test_det_sorted_lists_reuse_matches_plain(n, nv):
    """Factorisation reuse: sorted column lists (apyib_det_sort_lists) through the *_sorted entry points give
    the tables / products of the plain entry points, and both match numpy; generic O(1) overlap."""
    from apyib_b200._lib import lib, check
    from apyib_b200.aats import _Tables, _det_matvec, _det_outer
    from apyib_b200.device import to_device, to_host
    import apyib_b200
    rng = np.random.default_rng(300 + n)
    ns = n + nv
    S = np.eye(ns) + 0.3 * (rng.standard_normal((ns, ns)) + 1j * rng.standard_normal((ns, ns)))
    T = _Tables.get(n, 0, nv)
    dS = to_device(S, torch.complex128)
    assert T.LS[1] is not None and T.LS[2] is not None
    # the sort is a permutation of the enumeration and only moves substituted entries to the end
    cs, sg, ix = (x.cpu().numpy() for x in T.LS[2])
    assert sorted(ix.tolist()) == list(range(T.n2)) and set(np.unique(sg)) <= {-1.0, 1.0}
    assert (np.sort(cs, axis=1) == np.sort(T.L[2].cpu().numpy()[ix], axis=1)).all() and (cs[:, :n - 2] < n).all()
    for ck in (1, 2):
        rows, cols = T.L[2 if ck == 1 else 1], T.L[ck]
        want = np.linalg.det(S[rows.cpu().numpy()[:, None, :, None], cols.cpu().numpy()[None, :, None, :]])
        got_plain = to_host(_det_outer(dS, n, rows, cols))
        got_sorted = to_host(_det_outer(dS, n, rows, cols, T.LS[ck]))
        scale = max(1.0, np.abs(want).max())
        assert np.abs(got_plain - want).max() < 1e-11 * scale
        assert np.abs(got_sorted - want).max() < 1e-11 * scale
        Y = rng.standard_normal((3, cols.shape[0])) + 1j * rng.standard_normal((3, cols.shape[0]))
        Z = to_host(_det_matvec(dS, n, rows, cols, to_device(Y, torch.complex128), T.LS[ck]))
        assert np.abs(Z - Y @ want.T).max() < 1e-10 * max(1.0, np.abs(Y @ want.T).max())
    old = apyib_b200.config.LU_REUSE
    apyib_b200.config.LU_REUSE = False
    try:
        assert np.abs(to_host(_det_outer(dS, n, T.L[1], T.L[2], T.LS[2])) - to_host(_det_outer(dS, n, T.L[1], T.L[2]))).max() == 0
    finally:
        apyib_b200.config.LU_REUSE = old


def test_det_kernels_agree_on_fd_overlap():
    """Finite-difference-like overlap (I + O(h)): the thread-per-matrix and the sub-warp LU give the
    same doubles x doubles products to rounding (the O(h^2)..O(h^4) determinants of aats.py:581-618)."""
    from apyib_b200._lib import lib, check
    from apyib_b200.aats import _Tables, _det_matvec
    from apyib_b200.device import to_device, to_host
    no, nv = 6, 7
    rng = np.random.default_rng(77)
    S = np.eye(no + nv) + 1e-4 * (rng.standard_normal((no + nv,) * 2) + 0.1j * rng.standard_normal((no + nv,) * 2))
    T = _Tables.get(no, 0, nv)
    Y = to_device(rng.standard_normal((2, T.n2)) + 1j * rng.standard_normal((2, T.n2)), torch.complex128)
    dS = to_device(S, torch.complex128)
    res = []
    for which in (0, 1):
        check(lib.apyib_det_set_kernel(which))
        res.append(to_host(_det_matvec(dS, no, T.L[2], T.L[2], Y)))
    check(lib.apyib_det_set_kernel(0))
    scale = np.abs(res[1]).max()
    assert scale > 0 and np.abs(res[0] - res[1]).max() < 1e-9 * scale


@pytest.mark.parametrize("cplx", [False, True])
def test_dots_and_axpby(cplx):
    from apyib_b200._lib import lib, check
    from apyib_b200.device import to_device, to_host, zeros, ptr, stream_ptr, reduce_scratch, dtype_code
    rng = np.random.default_rng(6)
    n, nv = 100003, 5
    X, y = _rand(rng, (nv, n), cplx), _rand(rng, (n,), cplx)
    dX, dy = to_device(X), to_device(y)
    out = zeros((2 * nv,), torch.float64)
    for cj in (0, 1):
        check(lib.apyib_dots(dtype_code(dX), ptr(dX), n, nv, ptr(dy), n, cj, ptr(out), ptr(reduce_scratch()), stream_ptr()))
        got = to_host(out).reshape(nv, 2)
        want = (X.conj() if cj else X) @ y
        assert np.abs(got[:, 0] + 1j * got[:, 1] - want).max() < 1e-10
    check(lib.apyib_axpby(dtype_code(dX), n, 0.5, 0.25 if cplx else 0.0, ptr(dX[1]), 1, 2.0, 0.0, ptr(dy), stream_ptr()))
    a = (0.5 + 0.25j) if cplx else 0.5
    assert np.abs(to_host(dy) - (a * X[1].conj() + 2.0 * y)).max() < 1e-13


@pytest.mark.parametrize("cplx", [False, True])
def test_contract_split_k_dot_like_shapes(cplx):
    """tiny M x N with long K takes the split-K path (partials + fixed-order reduce)"""
    from apyib_b200.contraction import contract, _plan
    from apyib_b200.device import to_device, to_host
    rng = np.random.default_rng(9)
    X = _rand(rng, (3, 6, 6, 20, 20), cplx)
    Z = _rand(rng, (2, 6, 20, 6, 20), cplx)
    out0 = _rand(rng, (3, 2), cplx)
    dX, dZ, dO = to_device(X), to_device(Z), to_device(out0)
    assert _plan("xijab,qiajb->xq", dX, dZ, dO)[8] > 1, "expected split-K"
    contract("xijab,qiajb->xq", dX, dZ, dO, 0.5, 2.0)
    want = 0.5 * np.einsum("xijab,qiajb->xq", X, Z) + 2.0 * out0
    assert np.abs(to_host(dO) - want).max() < 1e-11 * max(1.0, np.abs(want).max())
    v = _rand(rng, (70001,), cplx)
    w = _rand(rng, (70001,), cplx)
    dv = to_device(v)
    s = torch.zeros((), dtype=dv.dtype, device=dv.device)
    contract("r,r->", dv, to_device(w), s, 1.0, 0.0, conj_a=True)
    assert abs(to_host(s) - np.vdot(v, w)) < 1e-10


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(640, 1500, 333), (4900, 144, 4900), (1024, 1024, 40)])
def test_contract_tma_path_matches_einsum(cplx, shape):
    """k-contiguous 2-D operands go through the TMA-fed kernel (cp.async.bulk.tensor + mbarrier ring);
    odd K exercises the hardware zero fill of the K tail, M/N tails the row clipping"""
    import apyib_b200
    from apyib_b200.contraction import contract, _plan
    from apyib_b200.device import to_device, to_host
    M, N, K = shape
    rng = np.random.default_rng(21)
    if cplx and K % 1:
        pytest.skip("n/a")
    A, B = _rand(rng, (M, K), cplx), _rand(rng, (N, K), cplx)
    out0 = _rand(rng, (M, N), cplx)
    dA, dB, dO = to_device(A), to_device(B), to_device(out0)
    plan = _plan("mk,nk->mn", dA, dB, dO)
    if (not cplx) and K % 2:
        assert True   # real operands with odd leading dimension are not 16-byte aligned per row -> gather kernel
    else:
        assert plan[10] is not None, "expected the TMA path"
    contract("mk,nk->mn", dA, dB, dO, 0.5, 1.5, conj_b=cplx)
    want = 0.5 * A @ (B.conj() if cplx else B).T + 1.5 * out0
    assert np.abs(to_host(dO) - want).max() < 1e-11 * max(1.0, np.abs(want).max())
    # same result from the gather kernel
    apyib_b200.config.USE_TMA = False
    try:
        dO2 = to_device(out0)
        contract("mk,nk->mn", dA, dB, dO2, 0.5, 1.5, conj_b=cplx)
    finally:
        apyib_b200.config.USE_TMA = True
    assert np.abs(to_host(dO2) - to_host(dO)).max() < 1e-11 * max(1.0, np.abs(want).max())


def test_contract_tma_batched_ladder():
    """batched ladder term (leading finite-difference-point index) through the 3-D tensor maps"""
    from apyib_b200.contraction import contract, _plan
    from apyib_b200.device import to_device, to_host
    rng = np.random.default_rng(22)
    nb, o, v = 40, 6, 14
    W = _rand(rng, (nb, v, v, v, v), True)
    t2 = _rand(rng, (nb, o, o, v, v), True)
    r = _rand(rng, (nb, o, o, v, v), True)
    dW, dt, dr = to_device(W), to_device(t2), to_device(r)
    assert _plan("sabcd,sijcd->sijab", dW, dt, dr)[10] is not None
    contract("sabcd,sijcd->sijab", dW, dt, dr, 0.5, 1.0)
    want = r + 0.5 * np.einsum("sabcd,sijcd->sijab", W, t2)
    assert np.abs(to_host(dr) - want).max() < 1e-11 * np.abs(want).max()


@pytest.mark.parametrize("n,nv,nf", [(3, 4, 0), (4, 5, 1), (6, 3, 0), (7, 6, 2), (9, 5, 0), (9, 13, 0), (10, 4, 0), (12, 4, 1)])
@pytest.mark.parametrize("kind", ["generic", "fd", "big_S"])
def test_det_prefix_shared_lu_matches_per_matrix_lu(n, nv, nf, kind):
    """csrc/dets_pairs.cu: one pivoted LU of the unsubstituted columns per (row list, group) + Schur-complement
    vectors of the candidate columns  ==  an LU of every substituted matrix (numpy, and the thread-per-matrix
    kernel), for singly (k = 1) and doubly (k = 2) substituted column tables, O(1) random overlaps (heavy
    pivoting), finite-difference-like overlaps, and S too large for shared memory."""
    import apyib_b200
    from apyib_b200.aats import _Tables, _det_matvec
    from apyib_b200.device import to_device, to_host
    if kind == "big_S" and (n, nv) not in ((9, 5), (4, 5)):
        pytest.skip("large-S variant checked for two sizes")
    rng = np.random.default_rng(700 + 13 * n + nv)
    ns = n + nv
    pad = 110 if kind == "big_S" else 0               # ns^2 * 16 B > shared memory left over -> S through L1
    h = 1e-4 if kind == "fd" else 0.3
    S = np.eye(ns + pad) + h * (rng.standard_normal((ns + pad,) * 2) + 1j * rng.standard_normal((ns + pad,) * 2))
    T = _Tables.get(n, nf, nv)
    dS = to_device(S, torch.complex128)
    cfg = apyib_b200.config
    old = cfg.LU_PREFIX
    try:
        for ck in (1, 2):
            if T.L[ck].shape[0] == 0:
                continue
            assert T.PFX[ck] is not None, "sorted lists must have the group structure"
            gl, cand, nc = T.PFX[ck]
            assert nc == nv and gl == (nv if ck == 1 else nv * (nv - 1) // 2)
            for rk in (2, 1, 0):
                rows, cols = T.L[rk], T.L[ck]
                if rows.shape[0] == 0:
                    continue
                want_D = np.linalg.det(S[rows.cpu().numpy()[:, None, :, None], cols.cpu().numpy()[None, :, None, :]])
                for ny in (1, 3):
                    Y = rng.standard_normal((ny, cols.shape[0])) + 1j * rng.standard_normal((ny, cols.shape[0]))
                    dY = to_device(Y, torch.complex128)
                    want = Y @ want_D.T
                    cfg.LU_PREFIX = False
                    base = to_host(_det_matvec(dS, n, rows, cols, dY, T.LS[ck], T.PFX[ck], ck))
                    cfg.LU_PREFIX = True
                    got = to_host(_det_matvec(dS, n, rows, cols, dY, T.LS[ck], T.PFX[ck], ck))
                    scale = max(np.abs(want).max(), 1e-300)
                    tol = 1e-10 if kind != "fd" else 1e-9      # fd: O(h^k) determinants, relative to the largest
                    assert np.abs(got - want).max() < tol * scale, (ck, rk, ny)
                    assert np.abs(got - base).max() < tol * scale, (ck, rk, ny)
    finally:
        cfg.LU_PREFIX = old


@pytest.mark.parametrize("n,nv,nf", [(4, 5, 1), (9, 6, 0), (12, 4, 0), (14, 3, 0)])
def test_det_stack_entry_points_match_single_overlap_calls(n, nv, nf):
    """apyib_det_outer_stack / _matvec_stack / _matvec_pairs_stack (grid.y = overlap) == the single-overlap entry
    points applied to every overlap of the stack, with one shared Y and with one Y per overlap; n = 14 goes
    through the sub-warp kernel (host loop over the stack)."""
    import ctypes as C
    from apyib_b200._lib import lib, check
    from apyib_b200.aats import _Tables, _det_matvec, _det_outer
    from apyib_b200.device import to_device, to_host, empty, ptr, stream_ptr
    rng = np.random.default_rng(900 + n)
    ns, nS = n + nv, 5
    Ss = np.stack([np.eye(ns) + 0.3 * (rng.standard_normal((ns, ns)) + 1j * rng.standard_normal((ns, ns))) for _ in range(nS)])
    dS = to_device(Ss, torch.complex128)
    T = _Tables.get(n, nf, nv)
    null = C.c_void_p(0)
    for rk, ck in ((1, 1), (2, 1), (1, 2), (2, 2), (0, 2), (2, 0)):
        rows, cols = T.L[rk], T.L[ck]
        nrow, ncol = rows.shape[0], cols.shape[0]
        if nrow == 0 or ncol == 0:
            continue
        variants = [(ptr(cols), null, null, None)]
        if T.LS[ck] is not None:
            cs, sg, ix = T.LS[ck]
            variants.append((ptr(cs), ptr(sg), ptr(ix), T.LS[ck]))
        for cp, sp, ip, ls in variants:
            out = empty((nS, nrow, ncol), torch.complex128)
            check(lib.apyib_det_outer_stack(ptr(dS), nS, ns, n, ptr(rows), nrow, cp, sp, ip, ncol, ptr(out), stream_ptr()))
            want = np.stack([to_host(_det_outer(dS[s], n, rows, cols, ls)) for s in range(nS)])
            assert np.abs(to_host(out) - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
            for ny, per in ((2, False), (3, True)):
                Y = rng.standard_normal((nS, ny, ncol)) + 1j * rng.standard_normal((nS, ny, ncol))
                dY = to_device(Y if per else Y[0], torch.complex128)
                want = np.stack([to_host(_det_matvec(dS[s], n, rows, cols, dY[s] if per else dY, ls)) for s in range(nS)])
                Z = empty((nS, ny, nrow), torch.complex128)
                work = empty((nS * int(lib.apyib_det_matvec_work_len(nrow, ncol, ny, n)),), torch.complex128)
                check(lib.apyib_det_matvec_stack(ptr(dS), nS, ns, n, ptr(rows), nrow, cp, sp, ip, ncol, ptr(dY),
                                                 ny * ncol if per else 0, ny, ptr(Z), ptr(work), stream_ptr()))
                scale = max(1.0, np.abs(want).max())
                assert np.abs(to_host(Z) - want).max() <= 1e-12 * scale
                if ls is not None and ck in (1, 2) and T.PFX[ck] is not None:
                    gl, cand, nc = T.PFX[ck]
                    work = empty((nS * int(lib.apyib_det_matvec_pairs_work_len(nrow, ncol // gl, ny, n, ck, ns, nc)),), torch.complex128)
                    Z2 = empty((nS, ny, nrow), torch.complex128)
                    check(lib.apyib_det_matvec_pairs_stack(ptr(dS), nS, ns, n, ck, ptr(rows), nrow, cp, sp, ip, ncol, gl, ptr(cand),
                                                           nc, ptr(dY), ny * ncol if per else 0, ny, ptr(Z2), ptr(work), stream_ptr()))
                    assert np.abs(to_host(Z2) - want).max() <= 1e-10 * scale


@pytest.mark.parametrize("n,nv", [(4, 5), (9, 6), (9, 13), (12, 4)])
def test_det_pairs_single_vector_variant_matches_generic(n, nv):
    """apyib_det_set_pairs_variant(1): the single-vector specialisation of the prefix-shared LU kernel must give
    the results of the generic kernel (same arithmetic, fewer predicated issue slots) to the last bits."""
    import apyib_b200
    from apyib_b200._lib import lib, check
    from apyib_b200.aats import _Tables, _det_matvec
    from apyib_b200.device import to_device, to_host
    rng = np.random.default_rng(40 + n)
    ns = n + nv
    S = to_device(np.eye(ns) + 0.3 * (rng.standard_normal((ns, ns)) + 1j * rng.standard_normal((ns, ns))), torch.complex128)
    T = _Tables.get(n, 0, nv)
    rows, cols = T.L[2], T.L[2]
    Y = to_device(rng.standard_normal((1, cols.shape[0])) + 1j * rng.standard_normal((1, cols.shape[0])), torch.complex128)
    old = apyib_b200.config.LU_PREFIX
    apyib_b200.config.LU_PREFIX = True
    try:
        res = []
        for variant in (0, 1):
            check(lib.apyib_det_set_pairs_variant(variant))
            res.append(to_host(_det_matvec(S, n, rows, cols, Y, T.LS[2], T.PFX[2], 2)))
        # same arithmetic; the two instantiations may contract their FMAs differently (last-bit differences)
        assert np.abs(res[0] - res[1]).max() <= 1e-13 * np.abs(res[0]).max()
    finally:
        check(lib.apyib_det_set_pairs_variant(1 if apyib_b200.config.PAIRS_SINGLE_VECTOR else 0))
        apyib_b200.config.LU_PREFIX = old
