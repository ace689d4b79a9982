"""TEST INFRASTRUCTURE (oracle): streamed restatement of apyib/aats.py:646-1055 for sizes where the reference's
8-index determinant tensor (aats.py:575, 8 TB at (S)-methyloxirane/cc-pVDZ) cannot exist.

`spatial_aat_terms_streamed` evaluates exactly the sums of `AAT.compute_spatial_aats` (same amplitude scaling
aats.py:690-711, same nine I_xy terms aats.py:714-1047, same signs / N factors), but never builds a determinant
tensor: every tensor element the einsums touch is evaluated on demand with numpy.linalg.det on the substituted
ndocc x ndocc overlap (aats.py:581-618).  An element of an antisymmetrically completed tensor (aats.py:620-630)
with unrestricted indices equals the determinant with row i replaced by virtual row a, row j by virtual row b
(columns likewise) and is zero for i == j or a == b, so no index ordering has to be tracked.

The cost is (non-zero bra amplitudes) x (non-zero ket amplitudes) determinants per term, so it is used with SPARSE
synthetic amplitudes (`sparse_aat_inputs`): the GPU path does not exploit the sparsity -- it runs its full dense
machinery on the same inputs -- while this oracle only visits the support.  Pinned by
tests/test_oracle_golden.py against the dense oracle (oracle/apyib_oracle.spatial_aat_terms, itself pinned against
the unmodified reference) at sizes where both run.
"""
from __future__ import annotations

import numpy as np

from oracle import apyib_oracle as orc


def sparse_aat_inputs(method, nbf, ndocc, nfzc, natom, seed, h=1e-4, amp=0.05, nnz2=12, nnz1=10):
    """orc.synthetic_aat_inputs with amplitudes that are zero outside a small random support
    (nnz2 doubles + their (j,i,b,a) partners, nnz1 singles per finite-difference point)."""
    rng = np.random.default_rng(seed)
    A = orc.AATInputs(method, nbf, ndocc, nfzc, h, h)
    o, v = ndocc - nfzc, nbf - ndocc

    def ovl():
        return np.eye(nbf) + h * (rng.standard_normal((nbf, nbf)) + 0.1j * rng.standard_normal((nbf, nbf)))

    def amps(cplx):
        dt = np.complex128 if cplx else np.float64
        val = lambda: amp * rng.standard_normal() + (0.1j * amp * rng.standard_normal() if cplx else 0)
        t1 = np.zeros((o, v), dtype=dt)
        t2 = np.zeros((o, o, v, v), dtype=dt)
        for _ in range(nnz1):
            t1[rng.integers(o), rng.integers(v)] = val()
        for _ in range(nnz2):
            i, j, a, b = rng.integers(o), rng.integers(o), rng.integers(v), rng.integers(v)
            x = val()
            t2[i, j, a, b] = x
            t2[j, i, b, a] = x
        return [1, t1, t2] if method == "CISD" else [1, 0, t2]

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


class _Dets:
    """On-demand substituted determinants of one MO overlap (aats.py:581-618)."""

    def __init__(self, S, no, nf):
        self.S, self.no, self.nf = np.asarray(S, dtype=np.complex128), no, nf

    def _lists(self, subs):
        """subs: (n, k, 2) int array of (occupied index relative to the frozen core, virtual index); returns the
        index lists (n, no) and a validity mask (False where two substitutions hit the same occupied orbital or
        insert the same virtual twice: the completed tensors are zero there)."""
        subs = np.asarray(subs, dtype=np.int64)
        if len(subs) == 0:
            return np.zeros((0, self.no), dtype=np.int64), np.zeros(0, dtype=bool)
        subs = subs.reshape(len(subs), -1, 2)
        n, k = subs.shape[0], subs.shape[1]
        L = np.tile(np.arange(self.no), (n, 1))
        ok = np.ones(n, dtype=bool)
        for q in range(k):
            L[np.arange(n), subs[:, q, 0] + self.nf] = subs[:, q, 1] + self.no
            for p in range(q):
                ok &= (subs[:, q, 0] != subs[:, p, 0]) & (subs[:, q, 1] != subs[:, p, 1])
        return L, ok

    def table(self, row_subs, col_subs):
        """D[r, c] = det S[rows(r), cols(c)] for every pair"""
        R, okr = self._lists(row_subs)
        Cc, okc = self._lists(col_subs)
        out = np.zeros((len(R), len(Cc)), dtype=np.complex128)
        chunk = max(1, 100000 // max(len(Cc), 1))
        for s in range(0, len(R), chunk):
            M = self.S[R[s:s + chunk, None, :, None], Cc[None, :, None, :]]
            out[s:s + chunk] = np.linalg.det(M)
        return out * okr[:, None] * okc[None, :]


def _nz1(x):
    idx = np.argwhere(x != 0)
    return idx, x[tuple(idx.T)]


def _nz2(x):
    """non-zeros of x[i,j,a,b] as substitution pairs ((i,a),(j,b)) and values"""
    idx = np.argwhere(x != 0)
    subs = np.stack([idx[:, [0, 2]], idx[:, [1, 3]]], axis=1) if len(idx) else np.zeros((0, 2, 2), dtype=np.int64)
    return idx, subs, x[tuple(idx.T)]


def spatial_aat_terms_streamed(A, alpha, beta, normalization="full"):
    """The nine I_xy partial sums of aats.py:646-1047 (dict, complex, before Im/(4 hR hB)); CISD / CID / MP2."""
    method = A.method
    cisd = method == "CISD"
    no, nf = A.ndocc, A.nfzc
    if normalization == "intermediate":
        N = N_np = N_nn = N_mp = N_mn = 1
    else:                                                                      # aats.py:652-669
        N = orc._spatial_norm(A.unperturbed_T, cisd)
        N_np = orc._spatial_norm(A.nuc_pos_T[alpha], cisd)
        N_nn = orc._spatial_norm(A.nuc_neg_T[alpha], cisd)
        N_mp = orc._spatial_norm(A.mag_pos_T[beta], cisd)
        N_mn = orc._spatial_norm(A.mag_neg_T[beta], cisd)
    d2 = lambda S: np.linalg.det(np.asarray(S)[:no, :no]) ** 2
    I = dict.fromkeys(("00", "0D", "D0", "DD", "0S", "S0", "SS", "SD", "DS"), 0)
    I["00"] = (d2(A.overlap_pp[alpha][beta]) * N_np * N_mp - d2(A.overlap_pn[alpha][beta]) * N_np * N_mn
               - d2(A.overlap_np[alpha][beta]) * N_nn * N_mp + d2(A.overlap_nn[alpha][beta]) * N_nn * N_mn)   # :672-677
    if cisd:                                                                   # aats.py:690-711
        t1 = N * A.unperturbed_T[1]
        t1_dH = N_mp * A.mag_pos_T[beta][1] - N_mn * A.mag_neg_T[beta][1]
        t1_c = np.conj(t1)
        t1_dR = np.conj(N_np * A.nuc_pos_T[alpha][1] - N_nn * A.nuc_neg_T[alpha][1])
    t2 = N * A.unperturbed_T[2]
    t2_dH = N_mp * A.mag_pos_T[beta][2] - N_mn * A.mag_neg_T[beta][2]
    t2_c = np.conj(t2)
    t2_dR = np.conj(N_np * A.nuc_pos_T[alpha][2] - N_nn * A.nuc_neg_T[alpha][2])
    asw = lambda t: t - t.swapaxes(2, 3)
    none = np.zeros((1, 0, 2), dtype=np.int64)

    def block(S, sign, x1, x2, y1, y2, s0_N=None, os_N=None, d0=False, od=False):
        D = _Dets(S, no, nf)
        dS = D.table(none, none)[0, 0]
        ix2, sx2, vx2 = _nz2(x2)
        iy2, sy2, vy2 = _nz2(y2)
        _, sxa, vxa = _nz2(asw(x2))
        _, sya, vya = _nz2(asw(y2))
        s1 = lambda sub2, q: sub2[:, q:q + 1, :]                 # the (i,a) resp. (j,b) half of a doubles entry
        # x~2 . A_iajb  and  y~2 . B_kcld                                   (second terms of :732, :722, :727, :749)
        xA2 = vxa @ D.table(sxa, none)[:, 0]
        yB2 = D.table(none, sya)[0, :] @ vya
        if cisd:
            i1, v1x = _nz1(x1)
            k1, v1y = _nz1(y1)
            sx1, sy1 = i1.reshape(-1, 1, 2), k1.reshape(-1, 1, 2)
            xA1 = v1x @ D.table(sx1, none)[:, 0]                # x1 . A_ia
            yB1 = D.table(none, sy1)[0, :] @ v1y                # y1 . B_kc
            if s0_N is not None:
                I["S0"] += sign * 2 * xA1 * dS * s0_N                                               # :746
            if os_N is not None:
                I["0S"] += sign * 2 * yB1 * dS * os_N                                               # :816
            if d0:                                                                                  # :749-750
                Aia, Ajb = D.table(s1(sx2, 0), none)[:, 0], D.table(s1(sx2, 1), none)[:, 0]
                I["D0"] += sign * (0.5 * xA2 * dS + np.sum(vx2 * Aia * Ajb))
            if od:                                                                                  # :819-820
                Bkc, Bld = D.table(none, s1(sy2, 0))[0, :], D.table(none, s1(sy2, 1))[0, :]
                I["0D"] += sign * (0.5 * yB2 * dS + np.sum(vy2 * Bkc * Bld))
            I["SS"] += sign * (2 * (v1x @ D.table(sx1, sy1) @ v1y) * dS + 2 * xA1 * yB1)            # :718-719
            I["DS"] += sign * (0.5 * (vxa @ D.table(sxa, sy1) @ v1y) * dS + 0.5 * xA2 * yB1         # :722-724
                               + 2 * np.sum((vx2 * D.table(s1(sx2, 1), none)[:, 0]) @ (D.table(s1(sx2, 0), sy1) @ v1y)))
            I["SD"] += sign * (0.5 * (v1x @ D.table(sx1, sya) @ vya) * dS + 0.5 * xA1 * yB2         # :727-729
                               + 2 * np.sum((v1x @ D.table(sx1, s1(sy2, 0))) * vy2 * D.table(none, s1(sy2, 1))[0, :]))
        # I_DD, aats.py:732-737
        dd = (vxa @ D.table(sxa, sya) @ vya) * dS + xA2 * yB2
        dd += 4 * np.sum((vxa @ D.table(sxa, s1(sy2, 0))) * vy2 * D.table(none, s1(sy2, 1))[0, :])
        dd += 2 * np.sum((vx2 * D.table(s1(sx2, 1), none)[:, 0]) @ (D.table(s1(sx2, 0), sya) @ vya))
        dd += 2 * np.sum((vx2 * D.table(s1(sx2, 0), none)[:, 0]) @ (D.table(s1(sx2, 1), sya) @ vya))
        dd += 8 * np.einsum("x,y,xy,xy->", vx2, vy2, D.table(s1(sx2, 0), s1(sy2, 0)), D.table(s1(sx2, 1), s1(sy2, 1)))
        I["DD"] += sign * 0.125 * dd

    x1r, x1c, y1h, y1t = (t1_dR, t1_c, t1_dH, t1) if cisd else (None,) * 4
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


def compute_spatial_aats_streamed(A, alpha, beta, normalization="full"):
    I = spatial_aat_terms_streamed(A, alpha, beta, normalization)
    return (1 / (4 * A.nuc_pert_strength * A.mag_pert_strength)) * np.imag(sum(I.values()))
