"""Host-side inputs of the hot path (geometry, AO-integral providers, Hamiltonian, complex-HF SCF): pure numpy,
shared by the product, the oracle and bench.py's reference arm.  See hostinputs/chem.py."""
from .chem import *            # noqa: F401,F403
from .chem import Molecule, BasisHandle, provider_ao_overlap, SGaussianProvider, SyntheticProvider, Psi4Provider, select_provider, Hamiltonian, hf_wfn, JK_HOOK   # noqa: F401
