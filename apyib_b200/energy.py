"""Method dispatch -- drop-in for apyib/energy.py (energy.py:17-81, 84-150).

Same signatures and return lists.  SCF and AO integrals are host inputs (hostchem); the
correlated solve runs on the device through mp2_wfn / ci_wfn.
"""
from __future__ import annotations

from .hostchem import Hamiltonian, hf_wfn, provider_ao_overlap
from .mp2_wfn import mp2_wfn
from .ci_wfn import ci_wfn
from .utils import compute_phase


def _correlated(parameters, wfn, print_level):
    E, t0, t1, t2 = 0, 1, 0, 0
    m = parameters["method"]
    if m == "MP2":
        E, t2 = mp2_wfn(parameters, wfn).solve_MP2()
    elif m == "CID":
        E, t2 = ci_wfn(parameters, wfn).solve_CID(print_level)
    elif m == "CISD":
        E, t1, t2 = ci_wfn(parameters, wfn).solve_CISD(print_level)
    elif m == "MP2_SO":
        E, t2 = mp2_wfn(parameters, wfn).solve_MP2_SO()
    elif m == "CID_SO":
        E, t2 = ci_wfn(parameters, wfn).solve_CID_SO(print_level)
    elif m == "CISD_SO":
        E, t1, t2 = ci_wfn(parameters, wfn).solve_CISD_SO(print_level)
    return E, [t0, t1, t2]


def scf_point(parameters, unperturbed_basis=None, unperturbed_C=None, print_level=0):
    """Host part of one finite-difference point: AO integrals, SCF and (optionally) the MO phase
    fix of energy.py:104.  Returns the hf_wfn-like object the correlated solvers consume."""
    H = Hamiltonian(parameters)
    wfn = hf_wfn(H)
    wfn.solve_SCF(parameters, print_level)
    if unperturbed_basis is not None:
        wfn.C = compute_phase(wfn.ndocc, wfn.nbf, unperturbed_basis, unperturbed_C, H.basis_set, wfn.C,
                              ao_overlap=provider_ao_overlap(unperturbed_basis, H.basis_set))
    return wfn


def correlated_many(parameters, wfns, print_level=0):
    """Device part for a list of points: [(E, [t0, t1, t2]), ...] in input order.  CI methods
    are solved with shared launches (ci_wfn.solve_many); MP2 is one streaming kernel per point."""
    from .ci_wfn import solve_many
    m = parameters["method"]
    if m in ("CID", "CISD", "CID_SO", "CISD_SO"):
        res = solve_many(m, parameters, wfns, print_level)
        return [(r[0], [1, r[1], r[2]] if len(r) == 3 else [1, 0, r[1]]) for r in res]
    from .utils import ao_prefetch
    ao_prefetch(wfns)            # queue every point's AO-integral upload on the copy stream; point k+1 uploads while k transforms
    return [_correlated(parameters, w, print_level) for w in wfns]


def _report(parameters, E_SCF, E, E_nuc):
    print("Method: ", parameters["method"])
    print("Electronic Hartree-Fock Energy: ", E_SCF)
    if parameters["method"] != "RHF":
        print("Electronic Post-Hartree-Fock Energy: ", E)
    print("Total Energy: ", E_SCF + E + E_nuc)


def energy(parameters, return_H=False, print_level=0):
    H = Hamiltonian(parameters)
    wfn = hf_wfn(H)
    E_SCF, C = wfn.solve_SCF(parameters, print_level)
    basis, E_nuc = H.basis_set, H.E_nuc
    E, T_list = _correlated(parameters, wfn, print_level)
    E_list = [E_SCF, E, E_nuc]
    if print_level == 2:
        _report(parameters, E_SCF, E, E_nuc)
        print(parameters)
    if return_H:
        return E_list, T_list, C, basis, H
    return E_list, T_list, C, basis


def phase_corrected_energy(parameters, unperturbed_basis, unperturbed_C, return_H=False, print_level=0):
    H = Hamiltonian(parameters)
    wfn = hf_wfn(H)
    E_SCF, C = wfn.solve_SCF(parameters, print_level)
    basis, E_nuc = H.basis_set, H.E_nuc
    # energy.py:104: fix the MO phases against the unperturbed orbitals before the correlated solve
    wfn.C = compute_phase(wfn.ndocc, wfn.nbf, unperturbed_basis, unperturbed_C, basis, C,
                          ao_overlap=provider_ao_overlap(unperturbed_basis, basis))
    E, T_list = _correlated(parameters, wfn, print_level)
    E_list = [E_SCF, E, E_nuc]
    if print_level > 0:
        _report(parameters, E_SCF, E, E_nuc)
    if return_H:
        return E_list, T_list, wfn.C, basis, H
    return E_list, T_list, wfn.C, basis
