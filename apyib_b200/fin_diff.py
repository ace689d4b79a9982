"""Finite-difference driver -- drop-in for apyib/fin_diff.py (compute_AAT, compute_APT,
compute_Hessian, compute_Nuclear_Gradient, compute_Magnetic_Field_Gradient; fin_diff.py:12-510).

The reference walks the displacement / field points in one serial Python loop.  Here the list
of points is explicit (`aat_points`, `apt_points`), so that the same driver can (a) run them
all on one GPU or (b) take the share of one rank when the points are sharded over the GPUs of
a box (parallel.py); every point is a full, independent solve -- no data-path collective.
"""
from __future__ import annotations

import numpy as np

from .energy import scf_point, correlated_many
from .hostchem import Molecule
from .utils import release_ao


def aat_points(natom):
    """The 6N+6 points of compute_AAT in the reference's order (fin_diff.py:285-370)."""
    pts = [("R", a, +1) for a in range(3 * natom)] + [("R", a, -1) for a in range(3 * natom)]
    pts += [("B", b, +1) for b in range(3)] + [("B", b, -1) for b in range(3)]
    return pts


def point_cost(kind):
    """Relative cost used for the static partition.  A complex (field) solve is 4x the flops of a real one; measured
    on a B200 at the (S)-methyloxirane/cc-pVDZ shape it costs 5.3x in large batches (57 vs 10.7 ms per point) and
    more when a rank holds a single complex point (its launches cannot fill the device): 5.5."""
    return 5.5 if kind == "B" else 1.0


class finite_difference(object):
    def __init__(self, parameters, unperturbed_basis, unperturbed_C):
        self.parameters = parameters
        self.molecule = Molecule.from_string(self.parameters["geom"])
        self.geom = self.molecule.geometry()
        self.natom = self.molecule.natom()
        self.unperturbed_basis = unperturbed_basis
        self.unperturbed_C = unperturbed_C

    # -- helpers ------------------------------------------------------------------------------
    def _displaced(self, shifts):
        """parameters with the geometry shifted by {coordinate: delta}; the caller's dict is
        mutated and restored exactly like the reference does (fin_diff.py:292, 306-307)."""
        g = np.copy(self.geom)
        for a, d in shifts:
            g[a // 3][a % 3] += d
        self.molecule.set_geometry(g)
        return self.molecule.create_psi4_string_from_molecule()

    def _reset(self):
        self.molecule.set_geometry(self.geom)
        self.parameters["geom"] = self.molecule.create_psi4_string_from_molecule()

    def scf_aat_point(self, point, h_R, h_B):
        """Host SCF (+ MO phase fix) of one displaced / field point; `parameters` is mutated and
        restored like in the reference (fin_diff.py:292-307, 336-351)."""
        kind, idx, sign = point
        if kind == "R":
            self.parameters["geom"] = self._displaced([(idx, sign * h_R)])
            wfn = scf_point(self.parameters, self.unperturbed_basis, self.unperturbed_C)
            self._reset()
        else:
            self.parameters["F_mag"][idx] += sign * h_B
            wfn = scf_point(self.parameters, self.unperturbed_basis, self.unperturbed_C)
            self.parameters["F_mag"][idx] -= sign * h_B
        return wfn

    def solve_aat_point(self, point, h_R, h_B):
        wfn = self.scf_aat_point(point, h_R, h_B)
        E, T_list = correlated_many(self.parameters, [wfn])[0]
        return [wfn.E_SCF, E, wfn.H.E_nuc], T_list, wfn.C, wfn.H.basis_set

    # -- fin_diff.py:267-372 ------------------------------------------------------------------
    # AO integrals, MO integrals, integral blocks and DIIS history of a point live on the device while its batch is
    # being solved (~3 nbf^4 words); the AAT points are solved in chunks that stay below this budget.
    AAT_BATCH_BYTES = 120 << 30

    def compute_AAT(self, nuc_pert_strength, mag_pert_strength, points=None, unperturbed_wfn=None):
        """Returns the reference's 12-tuple of lists.  `points` (optional) restricts the work to a
        subset (sharding); entries not computed are left as None.  The points are processed in chunks bounded by
        AAT_BATCH_BYTES: host SCFs of a chunk, then its correlated solves on the GPU together (batched launches),
        then the chunk's AO integrals are released on both sides -- only (C, basis, T) survive per point, as in
        the reference's serial loop (fin_diff.py:285-370).  `unperturbed_wfn` (optional, sharded driver): an
        already converged SCF of the unperturbed point whose correlated solve joins the first chunk; the 12-tuple
        is then followed by its (E_corr, T_list)."""
        n3 = 3 * self.natom
        res = {("R", +1): ([None] * n3, [None] * n3, [None] * n3), ("R", -1): ([None] * n3, [None] * n3, [None] * n3),
               ("B", +1): ([None] * 3, [None] * 3, [None] * 3), ("B", -1): ([None] * 3, [None] * 3, [None] * 3)}
        pts = list(aat_points(self.natom) if points is None else points)
        self._point_energies = {}
        extra = None
        chunk_p, chunk_w, nbytes = [], [], 0

        def flush():
            nonlocal extra, nbytes
            wf = ([unperturbed_wfn] if (unperturbed_wfn is not None and extra is None) else []) + chunk_w
            if wf:
                solved = correlated_many(self.parameters, wf)
                if unperturbed_wfn is not None and extra is None:
                    extra, solved = solved[0], solved[1:]
                for pt, wfn, (E, T_list) in zip(chunk_p, chunk_w, solved):
                    Cs, Bs, Ts = res[(pt[0], pt[2])]
                    Cs[pt[1]], Bs[pt[1]], Ts[pt[1]] = wfn.C, wfn.H.basis_set, T_list
                    self._point_energies[pt] = wfn.E_SCF + E + wfn.H.E_nuc       # total energy (gradient drivers)
                    release_ao(wfn)
                    try:
                        wfn.H.ERI = None                       # nbf^4 words of host memory per point
                    except AttributeError:
                        pass
            chunk_p.clear()
            chunk_w.clear()
            nbytes = 0

        for pt in pts:
            w = self.scf_aat_point(pt, nuc_pert_strength, mag_pert_strength)
            chunk_p.append(pt)
            chunk_w.append(w)
            nbytes += 3 * (16 if np.iscomplexobj(w.C) else 8) * w.nbf ** 4
            if nbytes >= self.AAT_BATCH_BYTES:
                flush()
        flush()
        (npC, npB, npT), (nnC, nnB, nnT) = res[("R", +1)], res[("R", -1)]
        (mpC, mpB, mpT), (mnC, mnB, mnT) = res[("B", +1)], res[("B", -1)]
        out = (npC, nnC, npB, nnB, npT, nnT, mpC, mnC, mpB, mnB, mpT, mnT)
        return out + (extra,) if unperturbed_wfn is not None else out

    # -- energy-only double differences ---------------------------------------------------------
    # AO integrals + ERI_MO of a point live on the device while its batch is being solved; bound the batch.
    BATCH_BYTES = 24 << 30

    def _total_energies(self, settings):
        """Total energies E_SCF + E_corr + E_nuc of a list of points, each given as
        (geometry shifts [(coordinate, delta), ...], F_el increments [(axis, delta), ...]).
        What the reference does with one `energy(parameters)` call per point (fin_diff.py:53, 70, 174, ...)
        is split into the host SCFs of all points followed by batched device solves (shared launches;
        bit-identical to the point-by-point solves, tests/test_gpu_solvers.py)."""
        out = [None] * len(settings)
        wfns, idx = [], []

        def flush():
            if wfns:
                for k, w, (E, _) in zip(idx, wfns, correlated_many(self.parameters, wfns)):
                    out[k] = w.E_SCF + E + w.H.E_nuc
                    release_ao(w)
            wfns.clear()
            idx.clear()

        for k, (shifts, fields) in enumerate(settings):
            if shifts:
                self.parameters["geom"] = self._displaced(shifts)
            for b, d in fields:
                self.parameters["F_el"][b] += d
            w = scf_point(self.parameters)
            for b, d in fields:
                self.parameters["F_el"][b] -= d
            if shifts:
                self._reset()
            wfns.append(w)
            idx.append(k)
            if len(wfns) * 3 * 16 * w.nbf ** 4 >= self.BATCH_BYTES:
                flush()
        flush()
        return out

    # -- fin_diff.py:151-263 ------------------------------------------------------------------
    def apt_points(self):
        return [(a, sr, b, sf) for sr in (+1, -1) for a in range(3 * self.natom) for sf in (+1, -1) for b in range(3)]

    def solve_apt_point(self, point, h_R, h_F):
        return self.solve_apt_points([point], h_R, h_F)[0]

    def solve_apt_points(self, points, h_R, h_F):
        """Total energies of a subset of the 36N (R +- h_R, F +- h_F) points (sharding: parallel.py)."""
        return self._total_energies([([(a, sr * h_R)], [(b, sf * h_F)]) for a, sr, b, sf in points])

    def compute_APT(self, nuc_pert_strength, elec_pert_strength, energies=None):
        n3 = 3 * self.natom
        if energies is None:
            pts = self.apt_points()
            energies = dict(zip(pts, self.solve_apt_points(pts, nuc_pert_strength, elec_pert_strength)))
        mu = {}
        for sr in (+1, -1):
            mu[sr] = np.array([[-(energies[(a, sr, b, +1)] - energies[(a, sr, b, -1)]) / (2 * elec_pert_strength)
                                for b in range(3)] for a in range(n3)])
        return (mu[+1] - mu[-1]) / (2 * nuc_pert_strength)

    # -- fin_diff.py:27-147 -------------------------------------------------------------------
    def hessian_points(self):
        n3 = 3 * self.natom
        return [(a, sa, b, sb) for sa in (+1, -1) for a in range(n3) for sb in (+1, -1) for b in range(n3)]

    def compute_Hessian(self, nuc_pert_strength, energies=None):
        n3 = 3 * self.natom
        h = nuc_pert_strength
        if energies is None:
            pts = self.hessian_points()
            energies = dict(zip(pts, self._total_energies([([(a, sa * h), (b, sb * h)], []) for a, sa, b, sb in pts])))
        g = {sa: np.array([[(energies[(a, sa, b, +1)] - energies[(a, sa, b, -1)]) / (2 * h) for b in range(n3)]
                           for a in range(n3)]) for sa in (+1, -1)}
        return (g[+1] - g[-1]) / (2 * h)

    # -- fin_diff.py:376-447, 451-510 -----------------------------------------------------------
    def _gradient(self, kind, n, h_R, h_B):
        """Central-difference energy gradient over the R (n = 3N) or B (n = 3) points of compute_AAT,
        phase-corrected like them; returns (E+ - E-)/2h and the per-point C / basis / T lists."""
        pts = [(kind, i, +1) for i in range(n)] + [(kind, i, -1) for i in range(n)]
        # the same chunked point loop as compute_AAT (host SCFs, batched device solves, AO integrals released per
        # chunk); it also records the total energy of every point
        lists = self.compute_AAT(h_R, h_B, points=pts)
        E = [self._point_energies[pt] for pt in pts]
        h = h_R if kind == "R" else h_B
        grad = np.zeros(n)
        for i in range(n):
            grad[i] = np.real(E[i] - E[n + i]) / (2 * h)     # the reference stores into a float array
        k0 = 0 if kind == "R" else 6
        posC, negC, posB, negB, posT, negT = lists[k0:k0 + 6]
        return grad, posC, negC, posB, negB, posT, negT

    def compute_Nuclear_Gradient(self, nuc_pert_strength):
        """fin_diff.py:376-447: returns (gradient (N,3), nuc_pos_C, nuc_neg_C, nuc_pos_basis, nuc_neg_basis,
        nuc_pos_T, nuc_neg_T)."""
        out = self._gradient("R", 3 * self.natom, nuc_pert_strength, 0.0)
        return (out[0].reshape(self.natom, 3),) + out[1:]

    def compute_Magnetic_Field_Gradient(self, mag_pert_strength):
        """fin_diff.py:451-510: returns (gradient (3,), mag_pos_C, mag_neg_C, mag_pos_basis, mag_neg_basis,
        mag_pos_T, mag_neg_T)."""
        return self._gradient("B", 3, 0.0, mag_pert_strength)
