"""VCD hand-off -- the consumer side of the hot path (apyib/vcd.py:32-136).

`vcd(parameters).compute_vcd_from_input(Hessian, APT, AAT_elec)` takes the three tensors the drivers of this package
produce -- finite_difference.compute_Hessian -> (3N,3N), compute_APT / compute_parallel_apts -> (3N,3),
compute_parallel_aats -> (3N,3) -- and returns harmonic frequencies [cm^-1], IR intensities [km/mol] and VCD
rotational strengths [10^-44 esu^2 cm^2].  It is O((3N)^3) host post-processing (component 12: out of scope as
GPU work); it exists here so that BASELINE configs[2] ("MP2 AATs + VCD rotational strengths") runs end to end
without Psi4: the reference's version needs psi4.geometry / psi4.qcel for masses and constants, this one carries
them (most-abundant-isotope masses, CODATA 2014 as in qcelemental's default context).
"""
from __future__ import annotations

import numpy as np

from .hostchem import Molecule

# CODATA 2014 (vcd.py:37-42 reads them from psi4.qcel.constants)
_C = 299792458.0                 # speed of light in vacuum, m/s
_ME = 9.10938356e-31             # electron mass, kg
_NA = 6.022140857e+23            # Avogadro constant, 1/mol
_E = 1.6021766208e-19            # atomic unit of charge, C
_E0 = 8.854187817e-12            # electric constant, F/m
_H = 6.62607004e-34              # Planck constant, J s

# most abundant isotope masses [u] (what psi4.core.Molecule.mass returns by default)
MASS = {"H": 1.00782503223, "HE": 4.00260325413, "LI": 7.0160034366, "BE": 9.012183065, "B": 11.00930536,
        "C": 12.0, "N": 14.00307400443, "O": 15.99491461957, "F": 18.99840316273, "NE": 19.9924401762,
        "NA": 22.989769282, "MG": 23.985041697, "AL": 26.98153853, "SI": 27.97692653465, "P": 30.97376199842,
        "S": 31.9720711744, "CL": 34.968852682, "AR": 39.9623831237}


class vcd(object):
    def __init__(self, parameters):
        self.parameters = parameters
        self.molecule = Molecule.from_string(parameters["geom"])
        self.natom = self.molecule.natom()

    def _constants(self):
        hbar = _H / (2 * np.pi)
        ke = 1 / (4 * np.pi * _E0)
        alpha = ke * _E ** 2 / (hbar * _C)
        a0 = hbar / (_ME * _C * alpha)
        Eh = hbar ** 2 / (_ME * a0 ** 2)
        return dict(u=1 / (1000 * _NA),
                    freq=np.sqrt(Eh / (a0 * a0 * _ME)) / (2.0 * np.pi * _C * 100.0),          # au -> cm^-1
                    ir=(_E ** 2 * ke * _NA * np.pi) / (1000 * 3 * _ME * _C ** 2),             # au -> km/mol
                    vcd=(_E ** 2 * hbar * a0) / (_ME * _C) * (1000 * _C) ** 2 * 1e44)         # au -> 10^-44 esu^2 cm^2

    def normal_modes(self, Hessian):
        """mass-weighted Hessian -> (omega [au], S): the 3N-6 highest modes, Cartesian displacement per unit normal
        coordinate in the columns of S (vcd.py:74-96)"""
        k = self._constants()
        n3 = 3 * self.natom
        m = np.array([MASS[self.molecule.symbols[i // 3].upper()] * k["u"] / _ME for i in range(n3)])
        W = np.diag(1 / np.sqrt(m))
        lam, L = np.linalg.eigh(W @ np.asarray(Hessian) @ W)
        S = np.flip(W @ L, 1)[:, :n3 - 6]
        with np.errstate(invalid="ignore"):
            omega = np.sqrt(np.flip(lam)[:n3 - 6])          # imaginary modes of a non-stationary geometry -> nan, like the reference
        return omega, S

    def nuclear_aat(self):
        """J[lambda alpha, beta] = 1/4 sum_gamma eps_{alpha beta gamma} R_{lambda gamma} Z_lambda   (vcd.py:109-115)"""
        R = self.molecule.geometry()
        J = np.zeros((3 * self.natom, 3))
        for lam in range(self.natom):
            Z = self.molecule.true_atomic_number(lam)
            x, y, z = R[lam]
            J[3 * lam:3 * lam + 3] = 0.25 * Z * np.array([[0.0, z, -y], [-z, 0.0, x], [y, -x, 0.0]])
        return J

    def compute_vcd_from_input(self, Hessian, APT, AAT_elec, print_level=1):
        k = self._constants()
        omega, S = self.normal_modes(Hessian)
        P_i = np.asarray(APT).T @ S                                     # APT in the normal-mode basis
        M_i = (np.asarray(AAT_elec) + self.nuclear_aat()).T @ S         # electronic + nuclear AAT
        R = np.einsum("bi,bi->i", P_i.real, M_i.real)
        D = np.einsum("bi,bi->i", P_i.real, P_i.real)
        w, D, R = omega * k["freq"], D * k["ir"], R * k["vcd"]
        if print_level > 0:
            print("\nFrequency   IR Intensity   Rotational Strength")
            print(" (cm-1)      (km/mol)    (esu**2 cm**2 10**44)")
            print("----------------------------------------------")
            for i in range(len(w)):
                print(f" {w[i]:7.2f}     {D[i]:8.3f}        {R[i]:8.3f}")
        return w, D, R
