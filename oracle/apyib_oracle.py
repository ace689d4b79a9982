"""numpy restatement of apyib's correlated-wavefunction hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY -- never imported by the product package
(``apyib_b200``).  Each function cites the reference file:line it restates
(paths relative to /root/reference).  Parity status: PINNED -- every function
here is checked (tests/test_oracle_vs_reference_golden.py) against fixtures
under tests/golden/ that were produced by running the *unmodified* reference
classes (stub-imported, see oracle/ref_harness.py + tests/golden/make_golden.py)
and, for real molecules, against the hard-coded literals of the reference's own
test-suite (tests/golden/reference_literals.py).

The restatement is deliberately plain (np.einsum / np.linalg.det), vectorised
only where the reference's interpreted loops would make the CPU test-suite
unusably slow; the arithmetic and every index convention are the reference's.
"""
from __future__ import annotations

import itertools
import numpy as np

SPATIAL_METHODS = ("RHF", "MP2", "CID", "CISD")
SO_METHODS = ("MP2_SO", "CID_SO", "CISD_SO")


# ----------------------------------------------------------------------------
# duck-typed inputs (the `wfn` object the solvers read)
# ----------------------------------------------------------------------------
class _Basis:
    def __init__(self, nfzc, nbf):
        self._nfzc, self._nbf = nfzc, nbf

    def n_frozen_core(self):
        return self._nfzc

    def nbf(self):
        return self._nbf


class _Ham:
    def __init__(self, T, V, ERI, E_nuc, nfzc):
        self.T, self.V, self.ERI, self.E_nuc = T, V, ERI, E_nuc
        self.basis_set = _Basis(nfzc, T.shape[0])


class Wfn:
    """Minimal stand-in for apyib.hf_wfn.hf_wfn (hf_wfn.py:14-27)."""

    def __init__(self, C, eps, ndocc, T, V, ERI, E_SCF=0.0, E_nuc=0.0, nfzc=0):
        self.C, self.eps, self.ndocc = C, eps, ndocc
        self.nbf = C.shape[0]
        self.E_SCF = E_SCF
        self.H = _Ham(T, V, ERI, E_nuc, nfzc)


def synthetic_wfn(nbf, ndocc, seed, complex_=False, nfzc=0, scale=0.01):
    """Synthetic per-point inputs of SURVEY 8(d): Hermitian-symmetric (pq|rs),
    C = 1, T = diag(eps), V = 0."""
    rng = np.random.default_rng(seed)
    g = scale * rng.standard_normal((nbf,) * 4)
    if complex_:
        g = g + 0.1j * scale * rng.standard_normal((nbf,) * 4)
    g = g + g.transpose(2, 3, 0, 1)
    g = g + g.transpose(1, 0, 3, 2).conj()
    eps = np.sort(rng.standard_normal(nbf))
    eps[ndocc:] += 4.0
    T = np.diag(eps).astype(g.dtype)
    V = np.zeros_like(T)
    C = np.eye(nbf, dtype=g.dtype)
    return Wfn(C, eps, ndocc, T, V, g, 0.0, 0.0, nfzc)


def rotated_wfn(nbf, ndocc, seed, complex_=False, nfzc=0, scale=0.01):
    """Like synthetic_wfn but with a non-trivial (unitary) MO coefficient matrix,
    so that the AO->MO transform (utils.py:258-279) is actually exercised."""
    w = synthetic_wfn(nbf, ndocc, seed, complex_, nfzc, scale)
    rng = np.random.default_rng(seed + 77)
    A = rng.standard_normal((nbf, nbf))
    if complex_:
        A = A + 1j * rng.standard_normal((nbf, nbf))
    Q, _ = np.linalg.qr(A)
    Q = Q.astype(w.H.ERI.dtype)
    w.C = Q
    w.H.T = (Q * w.eps[None, :]) @ Q.conj().T          # C^H T C = diag(eps)
    return w


# ----------------------------------------------------------------------------
# a1  index tables                                         utils.py:184-213
# ----------------------------------------------------------------------------
def get_slices(parameters, wfn):
    nfzc = wfn.H.basis_set.n_frozen_core()
    nbf, no = wfn.nbf, wfn.ndocc
    C_list = [slice(0, nfzc), slice(nfzc, no), slice(no, nbf), slice(nfzc, nbf)]
    m = parameters["method"]
    if m in SPATIAL_METHODS:
        I_list = [slice(0, nfzc), slice(0, no - nfzc), slice(no - nfzc, nbf - nfzc), slice(0, nbf - nfzc)]
    elif m in SO_METHODS:
        I_list = [slice(0, 2 * nfzc), slice(0, 2 * no - 2 * nfzc),
                  slice(2 * no - 2 * nfzc, 2 * nbf - 2 * nfzc), slice(0, 2 * nbf - 2 * nfzc)]
    else:
        raise ValueError(m)
    return C_list, I_list


# ----------------------------------------------------------------------------
# a2/a3  AO->MO                                            utils.py:217-279
# ----------------------------------------------------------------------------
def compute_F_MO(parameters, wfn, C_list):
    f, o, v, t = C_list
    C = wfn.C
    h = wfn.H.T + wfn.H.V
    G = wfn.H.ERI
    GK = 2 * G - G.swapaxes(1, 2)
    E_fc = 0
    if parameters["freeze_core"] == True:  # noqa: E712  (reference semantics, utils.py:238)
        Cf = C[:, f]
        Dfc = np.einsum("mp,np->mn", Cf, Cf.conj())
        h_fc = h + np.einsum("ls,mnls->mn", Dfc, GK)
        E_fc = np.einsum("nm,mn->", Dfc, h + h_fc)
        h = h_fc
    D = np.einsum("mp,np->mn", C[:, o], C[:, o].conj())
    F_AO = h + np.einsum("ls,mnls->mn", D, GK)
    F_MO = np.einsum("ip,ij,jq->pq", C[:, t].conj(), F_AO, C[:, t], optimize=True)
    return F_MO, E_fc


def compute_ERI_MO(parameters, wfn, C_list):
    C = wfn.C[:, C_list[3]]
    X = np.einsum("mnlg,gs->mnls", wfn.H.ERI, C, optimize=True)
    X = np.einsum("mnls,lr->mnrs", X, C.conj(), optimize=True)
    X = np.einsum("nq,mnrs->mqrs", C, X, optimize=True)
    X = np.einsum("mp,mqrs->pqrs", C.conj(), X, optimize=True)
    return X


# ----------------------------------------------------------------------------
# a4  spin blocking                                        utils.py:283-365, 393-422
# ----------------------------------------------------------------------------
def spin_block_2(X):
    """X_SO[p,q] = X[p//2,q//2] * [p%2 == q%2]  (alpha = even, beta = odd)."""
    n0, n1 = X.shape
    Y = np.repeat(np.repeat(X, 2, 0), 2, 1)
    p = np.arange(2 * n0)[:, None] % 2
    q = np.arange(2 * n1)[None, :] % 2
    return Y * (p == q)


def spin_block_4(X):
    Y = X
    for ax in range(4):
        Y = np.repeat(Y, 2, ax)
    par = [np.arange(Y.shape[ax]) % 2 for ax in range(4)]
    m_pq = par[0][:, None] == par[1][None, :]
    m_rs = par[2][:, None] == par[3][None, :]
    return Y * (m_pq[:, :, None, None] & m_rs[None, None, :, :])


# ----------------------------------------------------------------------------
# a5  MP2                                                  mp2_wfn.py:21-87
# ----------------------------------------------------------------------------
def _denoms(eps_o, eps_v):
    D1 = eps_o[:, None] - eps_v[None, :]
    D2 = eps_o[:, None, None, None] + eps_o[None, :, None, None] - eps_v[None, None, :, None] - eps_v
    return D1, D2


def solve_MP2(parameters, wfn):
    C_list, I_list = get_slices(parameters, wfn)
    o_, v_ = I_list[1], I_list[2]
    _, D2 = _denoms(wfn.eps[C_list[1]], wfn.eps[C_list[2]])
    W = compute_ERI_MO(parameters, wfn, C_list).swapaxes(1, 2)      # <pq|rs>
    t2 = W.swapaxes(0, 2).swapaxes(1, 3)[o_, o_, v_, v_] / D2
    L = 2 * W[o_, o_, v_, v_] - W.swapaxes(2, 3)[o_, o_, v_, v_]
    return np.einsum("ijab,ijab->", L, t2), t2


def solve_MP2_SO(parameters, wfn):
    C_list, I_list = get_slices(parameters, wfn)
    o_, v_ = I_list[1], I_list[2]
    _, D2 = _denoms(np.repeat(wfn.eps[C_list[1]], 2), np.repeat(wfn.eps[C_list[2]], 2))
    W = spin_block_4(compute_ERI_MO(parameters, wfn, C_list)).swapaxes(1, 2)
    A = W - W.swapaxes(2, 3)                                        # <pq||rs>
    t2 = A.swapaxes(0, 2).swapaxes(1, 3)[o_, o_, v_, v_] / D2
    return 0.25 * np.einsum("ijab,ijab->", A[o_, o_, v_, v_], t2), t2


# ----------------------------------------------------------------------------
# a10  DIIS                                                utils.py:104-140
# ----------------------------------------------------------------------------
def solve_general_DIIS(res_vec, t_vec, e_iter, t_iter, iteration, max_DIIS=7):
    while e_iter.shape[1] > max_DIIS:
        e_iter = e_iter[:, 1:]
        t_iter = t_iter[:, 1:]
    if iteration != 1:
        e_iter = np.hstack((e_iter, res_vec[:, None]))
        t_iter = np.hstack((t_iter, t_vec[:, None]))
    m = e_iter.shape[1]
    B = np.zeros((m + 1, m + 1), dtype=np.result_type(e_iter.dtype, np.float64))
    B[:m, :m] = e_iter.conj().T @ e_iter
    B[-1, :] = -1
    B[:, -1] = -1
    B[-1, -1] = 0
    rhs = np.zeros(m + 1)
    rhs[-1] = -1
    c = np.linalg.solve(B, rhs)
    return t_iter @ c[:-1], e_iter, t_iter


class _Diis:
    def __init__(self, enabled):
        self.enabled, self.e, self.t = enabled, None, None

    def __call__(self, iteration, res_parts, t_parts):
        if not self.enabled:
            return t_parts
        r = np.concatenate([x.reshape(-1) for x in res_parts])
        t = np.concatenate([x.reshape(-1) for x in t_parts])
        if iteration == 1:
            self.e, self.t = r[:, None].copy(), t[:, None].copy()
        t, self.e, self.t = solve_general_DIIS(r, t, self.e, self.t, iteration)
        out, off = [], 0
        for x in t_parts:
            out.append(t[off:off + x.size].reshape(x.shape))
            off += x.size
        return out


def _converged(parameters, iteration, dE, rms_list):
    if iteration <= 1:
        return False
    ok = abs(dE) < parameters["e_convergence"]
    for r in rms_list:
        ok = ok and (r < parameters["d_convergence"])   # numpy complex '<' is lexicographic
    return bool(ok)


class _CI:
    """Integrals as the reference's ci_wfn.__init__ sets them up (ci_wfn.py:24-47)."""

    def __init__(self, parameters, wfn):
        self.parameters, self.wfn = parameters, wfn
        self.C_list, self.I_list = get_slices(parameters, wfn)
        self.eps_o, self.eps_v = wfn.eps[self.C_list[1]], wfn.eps[self.C_list[2]]
        self.D_ia, self.D_ijab = _denoms(self.eps_o, self.eps_v)
        self.F_MO, self.E_fc = compute_F_MO(parameters, wfn, self.C_list)
        self.ERI_MO = compute_ERI_MO(parameters, wfn, self.C_list)


# ----------------------------------------------------------------------------
# a7  spatial CID                                          ci_wfn.py:51-167
# ----------------------------------------------------------------------------
def solve_CID(parameters, wfn, return_iters=False):
    ci = _CI(parameters, wfn)
    o, v = ci.I_list[1], ci.I_list[2]
    F, W, D2 = ci.F_MO, ci.ERI_MO.swapaxes(1, 2), ci.D_ijab
    K = W.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v]               # [i,j,a,b] = <ab|ij>
    L = 2 * W[o, o, v, v] - W.swapaxes(2, 3)[o, o, v, v]
    t2 = K / D2
    E = np.einsum("ijab,ijab->", L, t2)
    diis = _Diis(parameters["DIIS"])
    it = 1
    while it <= parameters["max_iterations"]:
        E_old, t2_old = E, t2.copy()
        r = 0.5 * K
        r = r + np.einsum("ijae,be->ijab", t2, F[v, v])
        r -= np.einsum("imab,mj->ijab", t2, F[o, o])
        r += 0.5 * np.einsum("mnab,mnij->ijab", t2, W[o, o, o, o], optimize=True)
        r += 0.5 * np.einsum("ijef,abef->ijab", t2, W[v, v, v, v], optimize=True)
        r += np.einsum("imae,mbej->ijab", t2 - t2.swapaxes(2, 3), W[o, v, v, o], optimize=True)
        r += np.einsum("imae,mbej->ijab", t2, W[o, v, v, o] - W.swapaxes(2, 3)[o, v, v, o], optimize=True)
        r -= np.einsum("mjae,mbie->ijab", t2, W[o, v, o, v], optimize=True)
        r = r + r.swapaxes(0, 1).swapaxes(2, 3)
        r -= E * t2
        t2 = t2 + r / D2
        (t2,) = diis(it, [r], [t2])
        E = np.einsum("ijab,ijab->", L, t2)
        rms = np.sqrt(np.einsum("ijab,ijab->", t2_old - t2, t2_old - t2))
        if _converged(parameters, it, E_old - E, [rms]):
            break
        it += 1
    return (E, t2, min(it, parameters["max_iterations"])) if return_iters else (E, t2)


# ----------------------------------------------------------------------------
# a9  spin-orbital CID / CISD                              ci_wfn.py:171-416
# ----------------------------------------------------------------------------
def _so_setup(ci):
    eo, ev = np.repeat(ci.eps_o, 2), np.repeat(ci.eps_v, 2)
    D1, D2 = _denoms(eo, ev)
    F = spin_block_2(ci.F_MO)
    W = spin_block_4(ci.ERI_MO).swapaxes(1, 2)
    A = W - W.swapaxes(2, 3)                                       # <pq||rs>
    return D1, D2, F, A


def solve_CID_SO(parameters, wfn, return_iters=False):
    ci = _CI(parameters, wfn)
    o, v = ci.I_list[1], ci.I_list[2]
    _, D2, F, A = _so_setup(ci)
    K = A.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v]               # <ab||ij> at [i,j,a,b]
    t2 = K / D2
    E = 0.25 * np.einsum("ijab,ijab->", t2, A[o, o, v, v])
    diis = _Diis(parameters["DIIS"])
    Aovvo = A[o, v, v, o]
    it = 1
    while it <= parameters["max_iterations"]:
        E_old, t2_old = E, t2.copy()
        r = K.copy()
        r += np.einsum("ijae,be->ijab", t2, F[v, v]) + np.einsum("ijeb,ae->ijab", t2, F[v, v])
        r -= np.einsum("imab,mj->ijab", t2, F[o, o]) + np.einsum("mjab,mi->ijab", t2, F[o, o])
        r += 0.5 * np.einsum("mnab,mnij->ijab", t2, A[o, o, o, o], optimize=True)
        r += 0.5 * np.einsum("ijef,abef->ijab", t2, A[v, v, v, v], optimize=True)
        r += np.einsum("imae,mbej->ijab", t2, Aovvo, optimize=True)
        r += np.einsum("mjae,mbei->ijab", t2, Aovvo, optimize=True)
        r += np.einsum("imeb,maej->ijab", t2, Aovvo, optimize=True)
        r += np.einsum("mjeb,maei->ijab", t2, Aovvo, optimize=True)
        r -= E * t2
        t2 = t2 + r / D2
        (t2,) = diis(it, [r], [t2])
        E = 0.25 * np.einsum("ijab,ijab->", A[o, o, v, v], t2)
        rms = np.sqrt(np.einsum("ijab,ijab->", t2_old - t2, t2_old - t2))
        if _converged(parameters, it, E_old - E, [rms]):
            break
        it += 1
    return (E, t2, min(it, parameters["max_iterations"])) if return_iters else (E, t2)


def solve_CISD_SO(parameters, wfn, return_iters=False):
    ci = _CI(parameters, wfn)
    o, v = ci.I_list[1], ci.I_list[2]
    D1, D2, F, A = _so_setup(ci)
    K = A.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v]
    t1 = F.swapaxes(0, 1)[o, v] / D1
    t2 = K / D2
    energy = lambda t1, t2: (np.einsum("ia,ia->", t1, F[o, v])
                             + 0.25 * np.einsum("ijab,ijab->", t2, A[o, o, v, v]))
    E = energy(t1, t2)
    diis = _Diis(parameters["DIIS"])
    Aovvo = A[o, v, v, o]
    it = 1
    while it <= parameters["max_iterations"]:
        E_old, t1_old, t2_old = E, t1.copy(), t2.copy()
        r1 = F.swapaxes(0, 1)[o, v].copy()
        r1 -= np.einsum("ji,ja->ia", F[o, o], t1)
        r1 += np.einsum("ab,ib->ia", F[v, v], t1)
        r1 += np.einsum("jabi,jb->ia", Aovvo, t1)
        r1 += np.einsum("jb,ijab->ia", F[o, v], t2)
        r1 += 0.5 * np.einsum("ajcb,ijcb->ia", A[v, o, v, v], t2, optimize=True)
        r1 -= 0.5 * np.einsum("kjib,kjab->ia", A[o, o, o, v], t2, optimize=True)
        r1 -= E * t1
        r2 = K.copy()
        r2 -= np.einsum("kbij,ka->ijab", A[o, v, o, o], t1, optimize=True)
        r2 -= np.einsum("akij,kb->ijab", A[v, o, o, o], t1, optimize=True)
        r2 += np.einsum("abcj,ic->ijab", A[v, v, v, o], t1, optimize=True)
        r2 += np.einsum("abic,jc->ijab", A[v, v, o, v], t1, optimize=True)
        r2 += np.einsum("bc,ijac->ijab", F[v, v], t2)
        r2 += np.einsum("ac,ijcb->ijab", F[v, v], t2)
        r2 -= np.einsum("kj,ikab->ijab", F[o, o], t2)
        r2 -= np.einsum("ki,kjab->ijab", F[o, o], t2)
        r2 += 0.5 * np.einsum("klij,klab->ijab", A[o, o, o, o], t2, optimize=True)
        r2 += 0.5 * np.einsum("abcd,ijcd->ijab", A[v, v, v, v], t2, optimize=True)
        r2 += np.einsum("kbcj,ikac->ijab", Aovvo, t2, optimize=True)
        r2 += np.einsum("kbci,kjac->ijab", Aovvo, t2, optimize=True)
        r2 += np.einsum("kacj,ikcb->ijab", Aovvo, t2, optimize=True)
        r2 += np.einsum("kaci,kjcb->ijab", Aovvo, t2, optimize=True)
        r2 -= E * t2
        t1 = t1 + r1 / D1
        t2 = t2 + r2 / D2
        t1, t2 = diis(it, [r1, r2], [t1, t2])
        E = energy(t1, t2)
        rms1 = np.sqrt(np.einsum("ia,ia->", t1_old - t1, t1_old - t1))
        rms2 = np.sqrt(np.einsum("ijab,ijab->", t2_old - t2, t2_old - t2))
        if _converged(parameters, it, E_old - E, [rms1, rms2]):
            break
        it += 1
    return (E, t1, t2, min(it, parameters["max_iterations"])) if return_iters else (E, t1, t2)


# ----------------------------------------------------------------------------
# a8  spatial CISD                                         ci_wfn.py:420-574
# ----------------------------------------------------------------------------
def solve_CISD(parameters, wfn, return_iters=False):
    ci = _CI(parameters, wfn)
    o, v = ci.I_list[1], ci.I_list[2]
    F, W, D1, D2 = ci.F_MO, ci.ERI_MO.swapaxes(1, 2), ci.D_ia, ci.D_ijab
    K = W.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v]
    L = 2.0 * W[o, o, v, v] - W.swapaxes(2, 3)[o, o, v, v]
    t1 = F.swapaxes(0, 1)[o, v] / D1
    t2 = K / D2
    energy = lambda t1, t2: 2.0 * np.einsum("ia,ia->", t1, F[o, v]) + np.einsum("ijab,ijab->", t2, L)
    E = energy(t1, t2)
    diis = _Diis(parameters["DIIS"])
    Wovvo, Wovov = W[o, v, v, o], W[o, v, o, v]
    Lovvo = 2.0 * Wovvo - W.swapaxes(2, 3)[o, v, v, o]
    it = 1
    while it <= parameters["max_iterations"]:
        E_old, t1_old, t2_old = E, t1.copy(), t2.copy()
        r1 = F.swapaxes(0, 1)[o, v].copy()
        r1 -= np.einsum("ji,ja->ia", F[o, o], t1)
        r1 += np.einsum("ab,ib->ia", F[v, v], t1)
        r1 += np.einsum("jabi,jb->ia", Lovvo, t1)
        r1 += np.einsum("jb,ijab->ia", F[o, v], 2.0 * t2 - t2.swapaxes(2, 3))
        r1 += np.einsum("ajbc,ijbc->ia", 2.0 * W[v, o, v, v] - W.swapaxes(2, 3)[v, o, v, v], t2, optimize=True)
        r1 -= np.einsum("kjib,kjab->ia", 2.0 * W[o, o, o, v] - W.swapaxes(2, 3)[o, o, o, v], t2, optimize=True)
        r1 -= E * t1
        r2 = K.copy()
        r2 += np.einsum("abcj,ic->ijab", W[v, v, v, o], t1, optimize=True)
        r2 += np.einsum("abic,jc->ijab", W[v, v, o, v], t1, optimize=True)
        r2 -= np.einsum("kbij,ka->ijab", W[o, v, o, o], t1, optimize=True)
        r2 -= np.einsum("akij,kb->ijab", W[v, o, o, o], t1, optimize=True)
        r2 += np.einsum("ac,ijcb->ijab", F[v, v], t2)
        r2 += np.einsum("bc,ijac->ijab", F[v, v], t2)
        r2 -= np.einsum("ki,kjab->ijab", F[o, o], t2)
        r2 -= np.einsum("kj,ikab->ijab", F[o, o], t2)
        r2 += np.einsum("klij,klab->ijab", W[o, o, o, o], t2, optimize=True)
        r2 += np.einsum("abcd,ijcd->ijab", W[v, v, v, v], t2, optimize=True)
        r2 -= np.einsum("kbcj,ikca->ijab", Wovvo, t2, optimize=True)
        r2 += np.einsum("kaci,kjcb->ijab", Lovvo, t2, optimize=True)
        r2 -= np.einsum("kbic,kjac->ijab", Wovov, t2, optimize=True)
        r2 -= np.einsum("kaci,kjbc->ijab", Wovvo, t2, optimize=True)
        r2 += np.einsum("kbcj,ikac->ijab", Lovvo, t2, optimize=True)
        r2 -= np.einsum("kajc,ikcb->ijab", Wovov, t2, optimize=True)
        r2 -= E * t2
        t1 = t1 + r1 / D1
        t2 = t2 + r2 / D2
        t1, t2 = diis(it, [r1, r2], [t1, t2])
        E = energy(t1, t2)
        rms1 = np.sqrt(np.einsum("ia,ia->", t1_old - t1, t1_old - t1))
        rms2 = np.sqrt(np.einsum("ijab,ijab->", t2_old - t2, t2_old - t2))
        if _converged(parameters, it, E_old - E, [rms1, rms2]):
            break
        it += 1
    return (E, t1, t2, min(it, parameters["max_iterations"])) if return_iters else (E, t1, t2)


# ----------------------------------------------------------------------------
# a14  MO / SO overlaps                                    utils.py:370-422
# ----------------------------------------------------------------------------
def mo_overlap(C_bra, S_ao, C_ket):
    return np.einsum("mp,mn,nq->pq", C_bra.conj(), S_ao, C_ket, optimize=True)


# ----------------------------------------------------------------------------
# a15  spin-orbital substituted determinant                aats.py:120-130
# ----------------------------------------------------------------------------
def _swap_perm(n, pairs):
    p = np.arange(n)
    for x in range(0, len(pairs), 2):
        a, b = pairs[x], pairs[x + 1]
        p[a], p[b] = p[b], p[a]
    return p


def compute_SO_det(overlap, nocc, bra_indices, ket_indices):
    n = overlap.shape[0]
    r = _swap_perm(n, bra_indices)[:nocc]
    c = _swap_perm(n, ket_indices)[:nocc]
    return np.linalg.det(overlap[np.ix_(r, c)])


# ----------------------------------------------------------------------------
# a18  all substituted determinants of one MO overlap      aats.py:558-642
# ----------------------------------------------------------------------------
def det_index_tables(no, nf, nv):
    """Enumeration of aats.py:581-618 as integer tables (bit-exact contract).

    singles : (ns, 2)  rows (i, a)               i in [nf,no), a in [0,nv)
    doubles : (nd, 4)  rows (i, a, j, b)         i<j, a<b    (loop order i,a,j,b)
    Indices are *full-space* occupied indices (i includes the frozen offset)
    and zero-based virtual indices (full row = a + no)."""
    singles = np.array([(i, a) for i in range(nf, no) for a in range(nv)], dtype=np.int32).reshape(-1, 2)
    doubles = np.array([(i, a, j, b) for i in range(nf, no) for a in range(nv)
                        for j in range(i + 1, no) for b in range(a + 1, nv)], dtype=np.int32).reshape(-1, 4)
    return singles, doubles


def _batched_sub_dets(S, no, row_subs, col_subs):
    """det of S[rows, cols] for every (row substitution, column substitution) pair.
    row_subs: (nr, kr, 2) int [(i,a)...]; col_subs: (nc, kc, 2)."""
    nr, nc = len(row_subs), len(col_subs)
    rows = np.tile(np.arange(no), (nr, 1))
    for q in range(row_subs.shape[1]):
        rows[np.arange(nr), row_subs[:, q, 0]] = row_subs[:, q, 1] + no
    cols = np.tile(np.arange(no), (nc, 1))
    for q in range(col_subs.shape[1]):
        cols[np.arange(nc), col_subs[:, q, 0]] = col_subs[:, q, 1] + no
    out = np.empty((nr, nc), dtype=np.complex128)
    chunk = max(1, 200000 // max(nc, 1))
    for s in range(0, nr, chunk):
        R = rows[s:s + chunk]
        M = S[R[:, None, :, None], cols[None, :, None, :]]
        out[s:s + chunk] = np.linalg.det(M)
    return out


def compute_all_dets(overlap, ndocc, nfzc, nbf):
    """The 9 objects of aats.py:642, same shapes/index order, built from batched
    determinants over the reference's restricted enumeration and then completed
    antisymmetrically exactly as aats.py:620-630."""
    no, nf, nv = ndocc, nfzc, nbf - ndocc
    o = no - nf
    S = np.asarray(overlap, dtype=np.complex128)
    sing, doub = det_index_tables(no, nf, nv)
    s_sub = sing.reshape(-1, 1, 2)
    d_sub = doub.reshape(-1, 2, 2)
    none = np.zeros((1, 0, 2), dtype=np.int32)
    det_S = np.linalg.det(S[:no, :no])

    si, sa = sing[:, 0] - nf, sing[:, 1]
    di, da, dj, db = doub[:, 0] - nf, doub[:, 1], doub[:, 2] - nf, doub[:, 3]

    ia_S = np.zeros((o, nv), dtype=np.complex128)
    S_kc = np.zeros((o, nv), dtype=np.complex128)
    ia_S[si, sa] = _batched_sub_dets(S, no, s_sub, none)[:, 0]
    S_kc[si, sa] = _batched_sub_dets(S, no, none, s_sub)[0, :]

    iajb_S = np.zeros((o, nv, o, nv), dtype=np.complex128)
    S_kcld = np.zeros((o, nv, o, nv), dtype=np.complex128)
    ia_S_kc = np.zeros((o, nv, o, nv), dtype=np.complex128)
    if len(doub):
        iajb_S[di, da, dj, db] = _batched_sub_dets(S, no, d_sub, none)[:, 0]
        S_kcld[di, da, dj, db] = _batched_sub_dets(S, no, none, d_sub)[0, :]
    ia_S_kc[si[:, None], sa[:, None], si[None, :], sa[None, :]] = _batched_sub_dets(S, no, s_sub, s_sub)

    iajb_S_kc = np.zeros((o, nv, o, nv, o, nv), dtype=np.complex128)
    ia_S_kcld = np.zeros((o, nv, o, nv, o, nv), dtype=np.complex128)
    iajb_S_kcld = np.zeros((o, nv) * 4, dtype=np.complex128)
    if len(doub):
        iajb_S_kc[di[:, None], da[:, None], dj[:, None], db[:, None], si[None, :], sa[None, :]] = \
            _batched_sub_dets(S, no, d_sub, s_sub)
        # stored [i][a][j][b][k][c]: COLUMNS (i,a),(j,b) substituted, ROW (k,c)   (aats.py:604-606)
        ia_S_kcld[di[None, :], da[None, :], dj[None, :], db[None, :], si[:, None], sa[:, None]] = \
            _batched_sub_dets(S, no, s_sub, d_sub)
        iajb_S_kcld[di[:, None], da[:, None], dj[:, None], db[:, None],
                    di[None, :], da[None, :], dj[None, :], db[None, :]] = _batched_sub_dets(S, no, d_sub, d_sub)

    def asym4(X):     # aats.py:620-623
        return X - X.swapaxes(0, 2) - X.swapaxes(1, 3) + X.swapaxes(0, 2).swapaxes(1, 3)

    iajb_S, S_kcld, iajb_S_kc, ia_S_kcld = asym4(iajb_S), asym4(S_kcld), asym4(iajb_S_kc), asym4(ia_S_kcld)
    ia_S_kcld = ia_S_kcld.swapaxes(0, 4).swapaxes(1, 5).swapaxes(2, 4).swapaxes(3, 5)       # aats.py:629
    X = iajb_S_kcld                                                                           # aats.py:630
    X = X - X.swapaxes(0, 2) - X.swapaxes(1, 3) + X.swapaxes(0, 2).swapaxes(1, 3)
    X = X - X.swapaxes(4, 6) - X.swapaxes(5, 7) + X.swapaxes(4, 6).swapaxes(5, 7)
    iajb_S_kcld = X
    return det_S, ia_S, S_kc, iajb_S, S_kcld, ia_S_kc, iajb_S_kc, ia_S_kcld, iajb_S_kcld


# ----------------------------------------------------------------------------
# AAT inputs container (what aats.AAT.__init__ leaves on `self`, aats.py:23-115)
# ----------------------------------------------------------------------------
class AATInputs:
    """overlap_uu, overlap_up[3], overlap_un[3], overlap_pu[3N], overlap_nu[3N],
    overlap_pp/pn/np/nn[3N][3]; T lists [t0, t1, t2] for unperturbed / nuc_pos[3N] /
    nuc_neg[3N] / mag_pos[3] / mag_neg[3]."""

    def __init__(self, method, nbf, ndocc, nfzc, h_R, h_B):
        self.method, self.nbf, self.ndocc, self.nfzc = method, nbf, ndocc, nfzc
        self.nuc_pert_strength, self.mag_pert_strength = h_R, h_B


def synthetic_aat_inputs(method, nbf, ndocc, nfzc, natom, seed, h=1e-4, amp=0.05):
    """Random (unphysical) overlaps S = 1 + h(N + 0.1j N) and amplitudes, SURVEY 8(d).
    For *_SO methods the overlaps are spin-blocked MO overlaps and amplitudes are
    antisymmetric spin-orbital tensors."""
    rng = np.random.default_rng(seed)
    so = method in SO_METHODS
    A = AATInputs(method, nbf, ndocc, nfzc, h, h)
    o, v = ndocc - nfzc, nbf - ndocc

    def ovl():
        S = np.eye(nbf) + h * (rng.standard_normal((nbf, nbf)) + 0.1j * rng.standard_normal((nbf, nbf)))
        return spin_block_2(S) if so else S

    def amps(cplx):
        O, V = (2 * o, 2 * v) if so else (o, v)
        def rnd(*s):
            x = amp * rng.standard_normal(s)
            return x + (0.1j * amp * rng.standard_normal(s) if cplx else 0)
        t1 = rnd(O, V)
        t2 = rnd(O, O, V, V)
        if so:
            t2 = t2 - t2.swapaxes(0, 1)
            t2 = t2 - t2.swapaxes(2, 3)
        else:
            t2 = t2 + t2.swapaxes(0, 1).swapaxes(2, 3)
        if method.startswith("CISD"):
            return [1, t1, t2]
        return [1, 0, t2]

    n3 = 3 * natom
    A.overlap_uu = ovl()
    A.overlap_up = [ovl() for _ in range(3)]
    A.overlap_un = [ovl() for _ in range(3)]
    A.overlap_pu = [ovl() for _ in range(n3)]
    A.overlap_nu = [ovl() for _ in range(n3)]
    for name in ("pp", "pn", "np", "nn"):
        setattr(A, "overlap_" + name, [[ovl() for _ in range(3)] for _ in range(n3)])
    A.unperturbed_T = amps(False)
    A.nuc_pos_T = [amps(False) for _ in range(n3)]
    A.nuc_neg_T = [amps(False) for _ in range(n3)]
    A.mag_pos_T = [amps(True) for _ in range(3)]
    A.mag_neg_T = [amps(True) for _ in range(3)]
    return A


# ----------------------------------------------------------------------------
# a16/a19  spatial AAT element                             aats.py:646-1055
# ----------------------------------------------------------------------------
def _spatial_norm(T, cisd):
    t2 = T[2]
    x = T[0] + (2 * np.einsum("ijab,ijab->", t2.conj(), t2) - np.einsum("ijab,ijba->", t2.conj(), t2))
    if cisd:
        x = x + 2 * np.einsum("ia,ia->", np.conj(T[1]), T[1])
    return 1 / np.sqrt(x)


def spatial_aat_terms(A, alpha, beta, normalization="full"):
    """Returns dict of the nine I_xy partial sums (complex, before Im/(4 hR hB))."""
    method = A.method
    cisd = method == "CISD"
    no = A.ndocc
    if method == "RHF" or normalization == "intermediate":
        N = N_np = N_nn = N_mp = N_mn = 1
    else:
        N = _spatial_norm(A.unperturbed_T, cisd)
        N_np = _spatial_norm(A.nuc_pos_T[alpha], cisd)
        N_nn = _spatial_norm(A.nuc_neg_T[alpha], cisd)
        N_mp = _spatial_norm(A.mag_pos_T[beta], cisd)
        N_mn = _spatial_norm(A.mag_neg_T[beta], cisd)

    d2 = lambda S: np.linalg.det(S[:no, :no]) ** 2
    I = dict.fromkeys(("00", "0D", "D0", "DD", "0S", "S0", "SS", "SD", "DS"), 0)
    I["00"] = (d2(A.overlap_pp[alpha][beta]) * N_np * N_mp - d2(A.overlap_pn[alpha][beta]) * N_np * N_mn
               - d2(A.overlap_np[alpha][beta]) * N_nn * N_mp + d2(A.overlap_nn[alpha][beta]) * N_nn * N_mn)
    if method == "RHF":
        return I

    if cisd:
        t1 = N * A.unperturbed_T[1]
        t1_dH = N_mp * A.mag_pos_T[beta][1] - N_mn * A.mag_neg_T[beta][1]
        t1_c = np.conj(t1)
        t1_dR = np.conj(N_np * A.nuc_pos_T[alpha][1] - N_nn * A.nuc_neg_T[alpha][1])
    t2 = N * A.unperturbed_T[2]
    t2_dH = N_mp * A.mag_pos_T[beta][2] - N_mn * A.mag_neg_T[beta][2]
    t2_c = np.conj(t2)
    t2_dR = np.conj(N_np * A.nuc_pos_T[alpha][2] - N_nn * A.nuc_neg_T[alpha][2])

    ein = lambda *a: np.einsum(*a, optimize=True)
    asw = lambda t: t - t.swapaxes(2, 3)

    def block(S, sign, x1, x2, y1, y2, s0_N=None, os_N=None, d0=False, od=False):
        dS, Aia, Bkc, Aiajb, Bkcld, Miakc, Miajbkc, Miakcld, M8 = compute_all_dets(S, A.ndocc, A.nfzc, A.nbf)
        xa, ya = asw(x2), asw(y2)
        if cisd:
            if s0_N is not None:
                I["S0"] += sign * 2 * ein("ia,ia->", x1, Aia) * dS * s0_N
            if os_N is not None:
                I["0S"] += sign * 2 * ein("kc,kc->", y1, Bkc) * dS * os_N
            if d0:
                I["D0"] += sign * (0.5 * ein("ijab,iajb->", xa, Aiajb) * dS
                                   + ein("jb,jb->", ein("ijab,ia->jb", x2, Aia), Aia))
            if od:
                I["0D"] += sign * (0.5 * ein("klcd,kcld->", ya, Bkcld) * dS
                                   + ein("ld,ld->", ein("klcd,kc->ld", y2, Bkc), Bkc))
            I["SS"] += sign * (2 * ein("ia,kc,iakc->", x1, y1, Miakc) * dS
                               + 2 * ein("ia,ia->", x1, Aia) * ein("kc,kc->", y1, Bkc))
            I["DS"] += sign * (0.5 * ein("ijab,kc,iajbkc->", xa, y1, Miajbkc) * dS
                               + 0.5 * ein("ijab,iajb->", xa, Aiajb) * ein("kc,kc->", y1, Bkc)
                               + 2 * ein("ijab,kc,iakc,jb->", x2, y1, Miakc, Aia))
            I["SD"] += sign * (0.5 * ein("ia,klcd,iakcld->", x1, ya, Miakcld) * dS
                               + 0.5 * ein("ia,ia->", x1, Aia) * ein("klcd,kcld->", ya, Bkcld)
                               + 2 * ein("ia,klcd,iakc,ld->", x1, y2, Miakc, Bkc))
        I["DD"] += sign * 0.125 * (
            ein("ijab,klcd,iajbkcld->", xa, ya, M8) * dS
            + ein("ijab,iajb->", xa, Aiajb) * ein("klcd,kcld->", ya, Bkcld)
            + 4 * ein("ijab,klcd,iajbkc,ld->", xa, y2, Miajbkc, Bkc)
            + 2 * ein("ijab,klcd,iakcld,jb->", x2, ya, Miakcld, Aia)
            + 2 * ein("ijab,klcd,ia,jbkcld->", x2, ya, Aia, Miakcld)
            + 8 * ein("ijab,klcd,iakc,jbld->", x2, y2, Miakc, Miakc))

    x1r = t1_dR if cisd else None
    x1c = t1_c if cisd else None
    y1h = t1_dH if cisd else None
    y1t = t1 if cisd else None
    block(A.overlap_uu, +1, x1r, t2_dR, y1h, t2_dH)
    block(A.overlap_up[beta], +1, x1r, t2_dR, y1t, t2, s0_N=N_mp, d0=True)
    block(A.overlap_un[beta], -1, x1r, t2_dR, y1t, t2, s0_N=N_mn, d0=True)
    block(A.overlap_pu[alpha], +1, x1c, t2_c, y1h, t2_dH, os_N=N_np, od=True)
    block(A.overlap_nu[alpha], -1, x1c, t2_c, y1h, t2_dH, os_N=N_nn, od=True)
    block(A.overlap_pp[alpha][beta], +1, x1c, t2_c, y1t, t2, s0_N=N_mp, os_N=N_np, d0=True, od=True)
    block(A.overlap_pn[alpha][beta], -1, x1c, t2_c, y1t, t2, s0_N=N_mn, os_N=N_np, d0=True, od=True)
    block(A.overlap_np[alpha][beta], -1, x1c, t2_c, y1t, t2, s0_N=N_mp, os_N=N_nn, d0=True, od=True)
    block(A.overlap_nn[alpha][beta], +1, x1c, t2_c, y1t, t2, s0_N=N_mn, os_N=N_nn, d0=True, od=True)
    return I


def compute_spatial_aats(A, alpha, beta, normalization="full"):
    I = spatial_aat_terms(A, alpha, beta, normalization)
    tot = sum(I.values())
    return (1 / (4 * A.nuc_pert_strength * A.mag_pert_strength)) * np.imag(tot)


# ----------------------------------------------------------------------------
# a16/a17  spin-orbital brute-force AAT element            aats.py:134-554
# ----------------------------------------------------------------------------
def _so_norm(T, cisd):
    x = T[0] ** 2 + 0.25 * np.einsum("ijab,ijab->", np.conj(T[2]), T[2])
    if cisd:
        x = x + np.einsum("ia,ia->", np.conj(T[1]), T[1])
    return 1 / np.sqrt(x)


def _so_det_table(S, nocc, nso, n_bra, n_ket):
    """det for every *unrestricted* bra/ket index tuple, in the reference's loop
    order (i,a[,j,b]) x (k,c[,l,d]); sequential-swap semantics of aats.py:120-130."""
    occ, vir = range(nocc), range(nocc, nso)
    def tuples(n):
        if n == 0:
            return [()]
        if n == 1:
            return [(i, a) for i in occ for a in vir]
        return [(i, a, j, b) for i in occ for a in vir for j in occ for b in vir]
    bra, ket = tuples(n_bra), tuples(n_ket)
    R = np.array([_swap_perm(nso, t)[:nocc] for t in bra])
    Cc = np.array([_swap_perm(nso, t)[:nocc] for t in ket])
    out = np.empty((len(bra), len(ket)), dtype=np.complex128)
    chunk = max(1, 100000 // len(ket))
    for s in range(0, len(bra), chunk):
        M = S[R[s:s + chunk, None, :, None], Cc[None, :, None, :]]
        out[s:s + chunk] = np.linalg.det(M)
    return out


def so_aat_terms(A, alpha, beta, normalization="full"):
    method = A.method
    cisd = method == "CISD_SO"
    nocc, nso = 2 * A.ndocc, 2 * A.nbf
    O, V = nocc, nso - nocc
    if method == "RHF" or normalization == "intermediate":
        N = N_np = N_nn = N_mp = N_mn = 1
    else:
        N = _so_norm(A.unperturbed_T, cisd)
        N_np = _so_norm(A.nuc_pos_T[alpha], cisd)
        N_nn = _so_norm(A.nuc_neg_T[alpha], cisd)
        N_mp = _so_norm(A.mag_pos_T[beta], cisd)
        N_mn = _so_norm(A.mag_neg_T[beta], cisd)

    ov = lambda S: spin_block_2(S) if method == "RHF" else S
    S_uu = None if method == "RHF" else A.overlap_uu
    S_pp, S_pn = ov(A.overlap_pp[alpha][beta]), ov(A.overlap_pn[alpha][beta])
    S_np, S_nn = ov(A.overlap_np[alpha][beta]), ov(A.overlap_nn[alpha][beta])

    def stencils(nb, nk):
        D = lambda S: _so_det_table(S, nocc, nso, nb, nk)
        st = {}
        st["pppp"] = (D(S_pp) * N_np * N_mp - D(S_pn) * N_np * N_mn - D(S_np) * N_nn * N_mp + D(S_nn) * N_nn * N_mn)
        if nk > 0 and method != "RHF":
            st["pu"] = D(A.overlap_pu[alpha]) * N_np * N - D(A.overlap_nu[alpha]) * N_nn * N
        if nb > 0 and method != "RHF":
            st["up"] = D(A.overlap_up[beta]) * N * N_mp - D(A.overlap_un[beta]) * N * N_mn
        if nb > 0 and nk > 0:
            st["uu"] = D(S_uu) * N * N
        return st

    I = dict.fromkeys(("00", "0D", "D0", "DD", "0S", "S0", "SS", "SD", "DS"), 0)
    I["00"] = stencils(0, 0)["pppp"][0, 0]
    if method == "RHF":
        return I

    # amplitude vectors in the loop order (i,a) / (i,a,j,b)
    def vec2(t):   # t[i,j,a,b] -> [(i,a,j,b)]
        return np.asarray(t).transpose(0, 2, 1, 3).reshape(-1)
    U, P, Ng, Mp, Mn = A.unperturbed_T, A.nuc_pos_T[alpha], A.nuc_neg_T[alpha], A.mag_pos_T[beta], A.mag_neg_T[beta]
    t2 = vec2(U[2]); t2c = np.conj(t2)
    t2_dH = vec2(Mp[2] - Mn[2]); t2_dR = np.conj(vec2(P[2] - Ng[2]))

    st = stencils(0, 2)
    I["0D"] = 0.25 * (t2_dH @ st["pu"][0] + t2 @ st["pppp"][0])
    st = stencils(2, 0)
    I["D0"] = 0.25 * (t2_dR @ st["up"][:, 0] + t2c @ st["pppp"][:, 0])
    st = stencils(2, 2)
    I["DD"] = 0.0625 * (t2_dR @ st["uu"] @ t2_dH + t2_dR @ st["up"] @ t2 + t2c @ st["pu"] @ t2_dH + t2c @ st["pppp"] @ t2)
    if cisd:
        t1 = np.asarray(U[1]).reshape(-1); t1c = np.conj(t1)
        t1_dH = (Mp[1] - Mn[1]).reshape(-1); t1_dR = np.conj((P[1] - Ng[1]).reshape(-1))
        st = stencils(0, 1)
        I["0S"] = t1_dH @ st["pu"][0] + t1 @ st["pppp"][0]
        st = stencils(1, 0)
        I["S0"] = t1_dR @ st["up"][:, 0] + t1c @ st["pppp"][:, 0]
        st = stencils(1, 1)
        I["SS"] = t1_dR @ st["uu"] @ t1_dH + t1_dR @ st["up"] @ t1 + t1c @ st["pu"] @ t1_dH + t1c @ st["pppp"] @ t1
        st = stencils(1, 2)
        I["SD"] = 0.25 * (t1_dR @ st["uu"] @ t2_dH + t1_dR @ st["up"] @ t2 + t1c @ st["pu"] @ t2_dH + t1c @ st["pppp"] @ t2)
        st = stencils(2, 1)
        I["DS"] = 0.25 * (t2_dR @ st["uu"] @ t1_dH + t2_dR @ st["up"] @ t1 + t2c @ st["pu"] @ t1_dH + t2c @ st["pppp"] @ t1)
    return I


def compute_SO_aats(A, alpha, beta, normalization="full"):
    """aats.py:520-554.  NOTE: the reference takes .imag of every term separately
    and sums the scaled reals; summing first is identical in exact arithmetic."""
    I = so_aat_terms(A, alpha, beta, normalization)
    k = 1 / (4 * A.nuc_pert_strength * A.mag_pert_strength)
    return sum(k * np.imag(x) for x in I.values())


# ----------------------------------------------------------------------------
# a21  perturbed-amplitude (linear-response) CISD iterations   analytic_aats.py:742-885, 994-1137
# ----------------------------------------------------------------------------
def _cisd_linear(F, W, o, v, t1, t2):
    """linear part of the spatial CISD residual (ci_wfn.py:458-482) for Fock F and <pq|rs> W"""
    Lovvo = 2.0 * W[o, v, v, o] - W.swapaxes(2, 3)[o, v, v, o]
    r1 = -np.einsum("ji,ja->ia", F[o, o], t1) + np.einsum("ab,ib->ia", F[v, v], t1)
    r1 = r1 + np.einsum("jabi,jb->ia", Lovvo, t1)
    r1 += np.einsum("jb,ijab->ia", F[o, v], 2.0 * t2 - t2.swapaxes(2, 3))
    r1 += np.einsum("ajbc,ijbc->ia", 2.0 * W[v, o, v, v] - W.swapaxes(2, 3)[v, o, v, v], t2, optimize=True)
    r1 -= np.einsum("kjib,kjab->ia", 2.0 * W[o, o, o, v] - W.swapaxes(2, 3)[o, o, o, v], t2, optimize=True)
    r2 = np.einsum("abcj,ic->ijab", W[v, v, v, o], t1, optimize=True)
    r2 = r2 + np.einsum("abic,jc->ijab", W[v, v, o, v], t1, optimize=True)
    r2 -= np.einsum("kbij,ka->ijab", W[o, v, o, o], t1, optimize=True)
    r2 -= np.einsum("akij,kb->ijab", W[v, o, o, o], t1, optimize=True)
    r2 += np.einsum("ac,ijcb->ijab", F[v, v], t2) + np.einsum("bc,ijac->ijab", F[v, v], t2)
    r2 -= np.einsum("ki,kjab->ijab", F[o, o], t2) + np.einsum("kj,ikab->ijab", F[o, o], t2)
    r2 += np.einsum("klij,klab->ijab", W[o, o, o, o], t2, optimize=True)
    r2 += np.einsum("abcd,ijcd->ijab", W[v, v, v, v], t2, optimize=True)
    r2 -= np.einsum("kbcj,ikca->ijab", W[o, v, v, o], t2, optimize=True)
    r2 += np.einsum("kaci,kjcb->ijab", Lovvo, t2, optimize=True)
    r2 -= np.einsum("kbic,kjac->ijab", W[o, v, o, v], t2, optimize=True)
    r2 -= np.einsum("kaci,kjbc->ijab", W[o, v, v, o], t2, optimize=True)
    r2 += np.einsum("kbcj,ikac->ijab", Lovvo, t2, optimize=True)
    r2 -= np.einsum("kajc,ikcb->ijab", W[o, v, o, v], t2, optimize=True)
    return r1, r2


def solve_perturbed_CISD(parameters, wfn, t1, t2, E_CISD, dF_MO, dERI_MO, dE_guess=0.0, return_iters=False):
    """dt/dlambda of the CISD amplitudes for perturbed MO integrals (dF_MO, chemists' dERI_MO)."""
    ci = _CI(parameters, wfn)
    o, v = ci.I_list[1], ci.I_list[2]
    F, W, D1, D2 = ci.F_MO, ci.ERI_MO.swapaxes(1, 2), ci.D_ia, ci.D_ijab
    dF, dW = dF_MO, dERI_MO.swapaxes(1, 2)
    L = 2.0 * W[o, o, v, v] - W.swapaxes(2, 3)[o, o, v, v]
    dL = 2.0 * dW[o, o, v, v] - dW.swapaxes(2, 3)[o, o, v, v]
    dK = dW.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v]
    p1, p2 = _cisd_linear(dF, dW, o, v, t1, t2)
    dt1 = (-dE_guess * t1 + p1) / D1                                        # :743-772
    dt2 = (-dE_guess * t2 + p2) / D2
    proj = lambda dt1, dt2: (2.0 * np.einsum("ia,ia->", t1, dF[o, v]) + np.einsum("ijab,ijab->", t2, dL)
                             + 2.0 * np.einsum("ia,ia->", dt1, F[o, v]) + np.einsum("ijab,ijab->", dt2, L))
    dE = proj(dt1, dt2)
    diis = _Diis(parameters["DIIS"])
    it = 1
    while it <= parameters["max_iterations"]:
        dE_old, o1, o2 = dE, dt1.copy(), dt2.copy()
        q1, q2 = _cisd_linear(F, W, o, v, dt1, dt2)
        r1 = dF.swapaxes(0, 1)[o, v] - dE * t1 + p1 - E_CISD * dt1 + q1       # :788-806
        r2 = dK - dE * t2 + p2 - E_CISD * dt2 + q2                           # :808-846
        dt1 = dt1 + r1 / D1
        dt2 = dt2 + r2 / D2
        dt1, dt2 = diis(it, [r1, r2], [dt1, dt2])
        dE = proj(dt1, dt2)
        rms1 = np.sqrt(np.einsum("ia,ia->", o1 - dt1, o1 - dt1))
        rms2 = np.sqrt(np.einsum("ijab,ijab->", o2 - dt2, o2 - dt2))
        if _converged(parameters, it, dE_old - dE, [rms1, rms2]):
            break
        it += 1
    return (dE, dt1, dt2, min(it, parameters["max_iterations"])) if return_iters else (dE, dt1, dt2)


def solve_perturbed_CID(parameters, wfn, t2, E_CID, dF_MO, dERI_MO, dE_guess=0.0, return_iters=False):
    """dt2/dlambda of the CID amplitudes, analytic_aats.py:1553-1649 (magnetic field) and :1754-1850
    (nuclear displacement): the doubles <- doubles terms of the CISD loop (:1588-1614), convergence on
    the energy derivative and rms(dt2) (:1636-1639)."""
    ci = _CI(parameters, wfn)
    o, v = ci.I_list[1], ci.I_list[2]
    F, W, D2 = ci.F_MO, ci.ERI_MO.swapaxes(1, 2), ci.D_ijab
    dF, dW = dF_MO, dERI_MO.swapaxes(1, 2)
    L = 2.0 * W[o, o, v, v] - W.swapaxes(2, 3)[o, o, v, v]
    dL = 2.0 * dW[o, o, v, v] - dW.swapaxes(2, 3)[o, o, v, v]
    dK = dW.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v]
    z1 = np.zeros(ci.D_ia.shape, dtype=t2.dtype)
    lin = lambda f, w, x: _cisd_linear(f, w, o, v, z1, x)[1]
    p2 = lin(dF, dW, t2)
    dt2 = (-dE_guess * t2 + p2) / D2                                        # :1553-1567
    proj = lambda dt2: np.einsum("ijab,ijab->", t2, dL) + np.einsum("ijab,ijab->", dt2, L)   # :1570-1571
    dE = proj(dt2)
    diis = _Diis(parameters["DIIS"])
    it = 1
    while it <= parameters["max_iterations"]:
        dE_old, o2 = dE, dt2.copy()
        r2 = dK - dE * t2 + p2 - E_CID * dt2 + lin(F, W, dt2)               # :1581-1614
        dt2 = dt2 + r2 / D2
        (dt2,) = diis(it, [r2], [dt2])
        dE = proj(dt2)
        rms2 = np.sqrt(np.einsum("ijab,ijab->", o2 - dt2, o2 - dt2))
        if _converged(parameters, it, dE_old - dE, [rms2]):
            break
        it += 1
    return (dE, dt2, min(it, parameters["max_iterations"])) if return_iters else (dE, dt2)


# ----------------------------------------------------------------------------
# a21  closed-form MP2 perturbed amplitudes and the perturbed-integral builds of the analytic route
#      analytic_aats.py:347-352, 446-451 (dt2) and :730-733, 977-981 (dERI).  PARITY UNPINNED against outputs of the
#      reference (its surrounding routine needs Psi4 derivative integrals); a line-by-line transcription.
# ----------------------------------------------------------------------------
def perturbed_MP2_t2(t2, dF, dW, D_ijab, O, V, kind):
    """dW: perturbed PHYSICISTS' integrals d<pq|rs> over the active MO space, dF the perturbed Fock matrix.
    kind "H": analytic_aats.py:347-352 (magnetic field); kind "R": :446-451 (nuclear displacement)."""
    o, v = slice(0, O), slice(O, O + V)
    if kind == "H":
        dt2 = dW.swapaxes(0, 2).swapaxes(1, 3)[o, o, v, v].copy()
        dt2 = dt2 + np.einsum("ac,ijcb->ijab", dF[v, v], t2)
        dt2 = dt2 + np.einsum("bc,ijac->ijab", dF[v, v], t2)
        dt2 = dt2 - np.einsum("ki,kjab->ijab", dF[o, o], t2)
        dt2 = dt2 - np.einsum("kj,ikab->ijab", dF[o, o], t2)
    else:
        dt2 = dW[o, o, v, v].copy()
        dt2 = dt2 - np.einsum("kjab,ik->ijab", t2, dF[o, o])
        dt2 = dt2 - np.einsum("ikab,kj->ijab", t2, dF[o, o])
        dt2 = dt2 + np.einsum("ijcb,ac->ijab", t2, dF[v, v])
        dt2 = dt2 + np.einsum("ijac,cb->ijab", t2, dF[v, v])
    return dt2 / D_ijab


def build_dERI(U, W_full, nf, kind, core=None):
    """U: (nbf, nbf) CPHF coefficients; W_full: physicists' MO integrals over ALL nbf orbitals; t = [nf, nbf).
    kind "H": analytic_aats.py:730-733; kind "R": :977-981 (+ the derivative-integral term `core`)."""
    nbf = U.shape[0]
    t = slice(nf, nbf)
    sgn = -1.0 if kind == "H" else 1.0
    d = np.einsum("tr,pqts->pqrs", U[:, t], W_full[t, t, :, t], optimize=True)
    d = d + np.einsum("ts,pqrt->pqrs", U[:, t], W_full[t, t, t, :], optimize=True)
    d = d + sgn * np.einsum("tp,tqrs->pqrs", U[:, t], W_full[:, t, t, t], optimize=True)
    d = d + sgn * np.einsum("tq,ptrs->pqrs", U[:, t], W_full[t, :, t, t], optimize=True)
    return d if core is None else core + d
