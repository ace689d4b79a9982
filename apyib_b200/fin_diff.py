"""Finite-difference driver -- drop-in for apyib/fin_diff.py (compute_AAT, compute_APT,
compute_Hessian; fin_diff.py:12-372).

The reference walks the displacement / field points in one serial Python loop.  Here the list
of points is explicit (`aat_points`, `apt_points`), so that the same driver can (a) run them
all on one GPU or (b) take the share of one rank when the points are sharded over the GPUs of
a box (parallel.py); every point is a full, independent solve -- no data-path collective.
"""
from __future__ import annotations

import copy

import numpy as np

from .energy import energy, phase_corrected_energy, scf_point, correlated_many
from .hostchem import Molecule


def aat_points(natom):
    """The 6N+6 points of compute_AAT in the reference's order (fin_diff.py:285-370)."""
    pts = [("R", a, +1) for a in range(3 * natom)] + [("R", a, -1) for a in range(3 * natom)]
    pts += [("B", b, +1) for b in range(3)] + [("B", b, -1) for b in range(3)]
    return pts


def point_cost(kind):
    """Relative cost used for the static partition: complex (field) solves ~4x a real one."""
    return 4.0 if kind == "B" else 1.0


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
    def compute_AAT(self, nuc_pert_strength, mag_pert_strength, points=None):
        """Returns the reference's 12-tuple of lists.  `points` (optional) restricts the work to a
        subset (sharding); entries not computed are left as None.  All host SCFs run first, then the
        correlated solves of all points go to the GPU together (batched launches)."""
        n3 = 3 * self.natom
        res = {("R", +1): ([None] * n3, [None] * n3, [None] * n3), ("R", -1): ([None] * n3, [None] * n3, [None] * n3),
               ("B", +1): ([None] * 3, [None] * 3, [None] * 3), ("B", -1): ([None] * 3, [None] * 3, [None] * 3)}
        pts = list(aat_points(self.natom) if points is None else points)
        wfns = [self.scf_aat_point(pt, nuc_pert_strength, mag_pert_strength) for pt in pts]
        solved = correlated_many(self.parameters, wfns) if pts else []
        for pt, wfn, (E, T_list) in zip(pts, wfns, solved):
            Cs, Bs, Ts = res[(pt[0], pt[2])]
            Cs[pt[1]], Bs[pt[1]], Ts[pt[1]] = wfn.C, wfn.H.basis_set, T_list
        (npC, npB, npT), (nnC, nnB, nnT) = res[("R", +1)], res[("R", -1)]
        (mpC, mpB, mpT), (mnC, mnB, mnT) = res[("B", +1)], res[("B", -1)]
        return npC, nnC, npB, nnB, npT, nnT, mpC, mnC, mpB, mnB, mpT, mnT

    # -- fin_diff.py:151-263 ------------------------------------------------------------------
    def apt_points(self):
        return [(a, sr, b, sf) for sr in (+1, -1) for a in range(3 * self.natom) for sf in (+1, -1) for b in range(3)]

    def solve_apt_point(self, point, h_R, h_F):
        a, sr, b, sf = point
        self.parameters["geom"] = self._displaced([(a, sr * h_R)])
        self.parameters["F_el"][b] += sf * h_F
        E_list, T_list, C, basis = energy(self.parameters)
        self.parameters["F_el"][b] -= sf * h_F
        self._reset()
        return E_list[0] + E_list[1] + E_list[2]

    def compute_APT(self, nuc_pert_strength, elec_pert_strength, energies=None):
        n3 = 3 * self.natom
        if energies is None:
            energies = {pt: self.solve_apt_point(pt, nuc_pert_strength, elec_pert_strength) for pt in self.apt_points()}
        mu = {}
        for sr in (+1, -1):
            mu[sr] = np.array([[-(energies[(a, sr, b, +1)] - energies[(a, sr, b, -1)]) / (2 * elec_pert_strength)
                                for b in range(3)] for a in range(n3)])
        return (mu[+1] - mu[-1]) / (2 * nuc_pert_strength)

    # -- fin_diff.py:27-147 -------------------------------------------------------------------
    def compute_Hessian(self, nuc_pert_strength):
        n3 = 3 * self.natom
        h = nuc_pert_strength
        g = {}
        for sa in (+1, -1):
            rows = []
            for a in range(n3):
                e = {}
                for sb in (+1, -1):
                    for b in range(n3):
                        self.parameters["geom"] = self._displaced([(a, sa * h), (b, sb * h)])
                        E_list, T_list, C, basis = energy(self.parameters)
                        e[(b, sb)] = E_list[0] + E_list[1] + E_list[2]
                rows.append([(e[(b, +1)] - e[(b, -1)]) / (2 * h) for b in range(n3)])
                self._reset()
            g[sa] = np.array(rows)
        return (g[+1] - g[-1]) / (2 * h)
