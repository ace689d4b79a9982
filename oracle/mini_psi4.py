"""A psi4 look-alike for s-type Gaussian basis sets -- just enough of the Psi4 API for the
UNMODIFIED reference (hamiltonian.py, hf_wfn.py, energy.py, fin_diff.py, aats.py, parallel.py)
to run on H/He molecules such as its own (H2)_2 test case.  TEST INFRASTRUCTURE ONLY.

Integrals come from hostinputs.SGaussianProvider (closed-form s-Gaussian formulas);
that engine is itself pinned by the reference's hard-coded (H2)_2 energies and AAT tensors
(tests/golden/reference_literals.py).
"""
from __future__ import annotations

import types

import numpy as np

import hostinputs as hc

_options = {"basis": "STO-3G", "freeze_core": False}
_MASS = {"H": 1.00782503223, "HE": 4.00260325413}
# CODATA 2014 (qcelemental's default context, which psi4.qcel.constants exposes) -- vcd.py:37-42
_CODATA2014 = {"speed of light in vacuum": 299792458.0, "electron mass": 9.10938356e-31,
               "Avogadro constant": 6.022140857e+23, "atomic unit of charge": 1.6021766208e-19,
               "electric constant": 8.854187817e-12, "Planck constant": 6.62607004e-34}


class _Mat:
    def __init__(self, a):
        self.np = np.asarray(a)

    @staticmethod
    def from_array(a):
        return _Mat(np.array(a, dtype=float))

    def __array__(self, dtype=None, copy=None):
        return self.np


class _Molecule:
    def __init__(self, mol):
        self._m = mol

    def natom(self):
        return self._m.natom()

    def geometry(self):
        return _Mat(self._m.geometry())

    def set_geometry(self, mat):
        self._m.set_geometry(np.asarray(mat.np if hasattr(mat, "np") else mat))

    def create_psi4_string_from_molecule(self):
        return self._m.create_psi4_string_from_molecule()

    def true_atomic_number(self, i):
        return self._m.true_atomic_number(i)

    def nuclear_repulsion_energy(self, field=(0.0, 0.0, 0.0)):
        return self._m.nuclear_repulsion_energy(field)

    def mass(self, i):                                   # most abundant isotope, as psi4 / qcelemental report it
        return _MASS[self._m.symbols[i].upper()]

    def to_arrays(self):                                 # (geom [bohr], mass, elem, Z, uniq) -- vcd.py:107
        Z = np.array([self._m.true_atomic_number(i) for i in range(self._m.natom())], dtype=float)
        mass = np.array([self.mass(i) for i in range(self._m.natom())])
        elem = np.array(self._m.symbols)
        return self._m.geometry(), mass, elem, Z, elem

    def fix_orientation(self, *_):
        pass

    def fix_com(self, *_):
        pass

    def update_geometry(self):
        pass


class _BasisSet:
    def __init__(self, handle):
        self.handle = handle

    @staticmethod
    def build(molecule):
        prov = hc.SGaussianProvider(_options["basis"])
        return _BasisSet(prov.basis(molecule._m))

    def nbf(self):
        return self.handle.nbf()

    def n_frozen_core(self):
        return 0

    def molecule(self):
        return _Molecule(self.handle.molecule)

    def __eq__(self, other):
        return self is other

    def __ne__(self, other):
        return self is not other

    __hash__ = object.__hash__


class _MintsHelper:
    def __init__(self, basis):
        self.b = basis
        self._ints = None

    def _i(self):
        if self._ints is None:
            self._ints = self.b.handle.provider.integrals(self.b.handle.molecule)
        return self._ints

    def ao_kinetic(self):
        return _Mat(self._i()["T"])

    def ao_potential(self):
        return _Mat(self._i()["V"])

    def ao_eri(self):
        return _Mat(self._i()["ERI"])

    def ao_overlap(self, b1=None, b2=None):
        if b1 is None:
            return _Mat(self._i()["S"])
        return _Mat(b1.handle.provider.ao_overlap(b1.handle, b2.handle))

    def ao_dipole(self):
        return [_Mat(x) for x in self._i()["dipole"]]

    def ao_angular_momentum(self):
        return [_Mat(x) for x in self._i()["angmom"]]


def as_module():
    psi4 = types.ModuleType("psi4")
    core = types.ModuleType("psi4.core")
    core.clean_options = lambda: None
    core.clean = lambda: None
    core.BasisSet = _BasisSet
    core.MintsHelper = _MintsHelper
    core.Matrix = _Mat
    core.Molecule = _Molecule
    psi4.core = core
    psi4.set_options = lambda d: _options.update({k.lower(): v for k, v in d.items()})
    psi4.geometry = lambda s: _Molecule(hc.Molecule.from_string(s))
    psi4.set_output_file = lambda *a, **k: None
    psi4.set_memory = lambda *a, **k: None
    qcel = types.ModuleType("psi4.qcel")
    qcel.constants = types.SimpleNamespace(get=lambda name: _CODATA2014[name])
    psi4.qcel = qcel
    return psi4
