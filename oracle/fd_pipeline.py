"""CPU restatement of the reference's end-to-end finite-difference AAT/APT pipeline
(energy.py:17-150, fin_diff.py:151-372, aats.py:23-115, parallel.py:16-48) on top of the
oracle solvers.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Host inputs (geometry, AO integrals, SCF) come from hostinputs -- the same host
code the product uses; everything the product does on the GPU is done here with
oracle.apyib_oracle (numpy).
"""
from __future__ import annotations

import numpy as np

import hostinputs as hc
from oracle import apyib_oracle as orc


def _solve(parameters, wfn):
    m = parameters["method"]
    E, t1, t2 = 0, 0, 0
    if m == "MP2":
        E, t2 = orc.solve_MP2(parameters, wfn)
    elif m == "MP2_SO":
        E, t2 = orc.solve_MP2_SO(parameters, wfn)
    elif m in ("CID", "CID_SO"):
        E, t2 = getattr(orc, "solve_" + m)(parameters, wfn)
    elif m in ("CISD", "CISD_SO"):
        E, t1, t2 = getattr(orc, "solve_" + m)(parameters, wfn)
    return E, [1, t1, t2]


def energy(parameters):
    H = hc.Hamiltonian(parameters)
    wfn = hc.hf_wfn(H)
    E_SCF, C = wfn.solve_SCF(parameters)
    E, T = _solve(parameters, wfn)
    return [E_SCF, E, H.E_nuc], T, C, H.basis_set, wfn


def compute_phase(unperturbed_basis, unperturbed_C, basis, C):          # utils.py:427-446
    S = orc.mo_overlap(unperturbed_C, hc.provider_ao_overlap(unperturbed_basis, basis), C)
    d = np.diagonal(S)
    return C * ((d / np.sqrt(d * np.conj(d))) ** -1)[None, :]


def phase_corrected_energy(parameters, unperturbed_basis, unperturbed_C):
    H = hc.Hamiltonian(parameters)
    wfn = hc.hf_wfn(H)
    E_SCF, C = wfn.solve_SCF(parameters)
    wfn.C = compute_phase(unperturbed_basis, unperturbed_C, H.basis_set, C)
    E, T = _solve(parameters, wfn)
    return [E_SCF, E, H.E_nuc], T, wfn.C, H.basis_set


def fd_points(parameters, basis0, C0, h_R, h_B):
    mol = hc.Molecule.from_string(parameters["geom"])
    geom0 = mol.geometry()
    saved = parameters["geom"]
    out = {}
    for sign in (+1, -1):
        for a in range(3 * mol.natom()):
            g = geom0.copy()
            g[a // 3][a % 3] += sign * h_R
            mol.set_geometry(g)
            parameters["geom"] = mol.create_psi4_string_from_molecule()
            out[("R", a, sign)] = phase_corrected_energy(parameters, basis0, C0)
    parameters["geom"] = saved
    for sign in (+1, -1):
        for b in range(3):
            parameters["F_mag"][b] += sign * h_B
            out[("B", b, sign)] = phase_corrected_energy(parameters, basis0, C0)
            parameters["F_mag"][b] -= sign * h_B
    return out


def build_aat_inputs(parameters, h_R, h_B):
    E_list, T0, C0, basis0, wfn = energy(parameters)
    pts = fd_points(parameters, basis0, C0, h_R, h_B)
    natom = hc.Molecule.from_string(parameters["geom"]).natom()
    n3 = 3 * natom
    method = parameters["method"]
    so = method in orc.SO_METHODS
    A = orc.AATInputs(method, wfn.nbf, wfn.ndocc, basis0.n_frozen_core(), h_R, h_B)

    def ovl(bb, Cb, kb, Ck):
        S = orc.mo_overlap(Cb, hc.provider_ao_overlap(bb, kb), Ck)
        return orc.spin_block_2(S) if so else S

    P = lambda k, i, s: pts[(k, i, s)]
    A.unperturbed_T = T0
    A.nuc_pos_T = [P("R", a, +1)[1] for a in range(n3)]
    A.nuc_neg_T = [P("R", a, -1)[1] for a in range(n3)]
    A.mag_pos_T = [P("B", b, +1)[1] for b in range(3)]
    A.mag_neg_T = [P("B", b, -1)[1] for b in range(3)]
    Cb = lambda k, i, s: (P(k, i, s)[3], P(k, i, s)[2])
    if method != "RHF":
        A.overlap_uu = ovl(basis0, C0, basis0, C0)
        A.overlap_up = [ovl(basis0, C0, *Cb("B", b, +1)) for b in range(3)]
        A.overlap_un = [ovl(basis0, C0, *Cb("B", b, -1)) for b in range(3)]
        A.overlap_pu = [ovl(*Cb("R", a, +1), basis0, C0) for a in range(n3)]
        A.overlap_nu = [ovl(*Cb("R", a, -1), basis0, C0) for a in range(n3)]
    for name, sr, sb in (("pp", 1, 1), ("pn", 1, -1), ("np", -1, 1), ("nn", -1, -1)):
        setattr(A, "overlap_" + name, [[ovl(*Cb("R", a, sr), *Cb("B", b, sb)) for b in range(3)] for a in range(n3)])
    return A, natom, E_list


def compute_parallel_aats(parameters, h_R, h_B, normalization="full", terms=False):
    A, natom, _ = build_aat_inputs(parameters, h_R, h_B)
    spatial = parameters["method"] in orc.SPATIAL_METHODS
    k = 1 / (4 * h_R * h_B)
    I = np.zeros((3 * natom, 3))
    T = {}
    for a in range(3 * natom):
        for b in range(3):
            t = (orc.spatial_aat_terms if spatial else orc.so_aat_terms)(A, a, b, normalization)
            for name, v in t.items():
                T.setdefault(name, np.zeros((3 * natom, 3)))[a, b] = k * np.imag(v)
            I[a, b] = sum(k * np.imag(v) for v in t.values())
    return (I, T) if terms else I


def compute_APT(parameters, h_R, h_F):                                   # fin_diff.py:151-263
    mol = hc.Molecule.from_string(parameters["geom"])
    geom0, saved = mol.geometry(), parameters["geom"]
    n3 = 3 * mol.natom()
    mu = {}
    for sr in (+1, -1):
        rows = []
        for a in range(n3):
            g = geom0.copy()
            g[a // 3][a % 3] += sr * h_R
            mol.set_geometry(g)
            parameters["geom"] = mol.create_psi4_string_from_molecule()
            e = {}
            for sf in (+1, -1):
                for b in range(3):
                    parameters["F_el"][b] += sf * h_F
                    E_list = energy(parameters)[0]
                    e[(b, sf)] = E_list[0] + E_list[1] + E_list[2]
                    parameters["F_el"][b] -= sf * h_F
            rows.append([-(e[(b, +1)] - e[(b, -1)]) / (2 * h_F) for b in range(3)])
        mu[sr] = np.array(rows)
    parameters["geom"] = saved
    return (mu[+1] - mu[-1]) / (2 * h_R)


def compute_Hessian(parameters, h):                                      # fin_diff.py:27-147
    mol = hc.Molecule.from_string(parameters["geom"])
    geom0, saved = mol.geometry(), parameters["geom"]
    n3 = 3 * mol.natom()
    g = {}
    for sa in (+1, -1):
        rows = []
        for a in range(n3):
            e = {}
            for sb in (+1, -1):
                for b in range(n3):
                    x = geom0.copy()
                    x[a // 3][a % 3] += sa * h
                    x[b // 3][b % 3] += sb * h
                    mol.set_geometry(x)
                    parameters["geom"] = mol.create_psi4_string_from_molecule()
                    E_list = energy(parameters)[0]
                    e[(b, sb)] = E_list[0] + E_list[1] + E_list[2]
            rows.append([(e[(b, +1)] - e[(b, -1)]) / (2 * h) for b in range(n3)])
        g[sa] = np.array(rows)
    parameters["geom"] = saved
    return (g[+1] - g[-1]) / (2 * h)


def compute_Nuclear_Gradient(parameters, basis0, C0, h):                 # fin_diff.py:376-447
    pts = fd_points(parameters, basis0, C0, h, 0.0)
    n3 = 3 * hc.Molecule.from_string(parameters["geom"]).natom()
    tot = lambda p: p[0][0] + p[0][1] + p[0][2]
    grad = np.zeros(n3)
    for a in range(n3):
        grad[a] = np.real(tot(pts[("R", a, +1)]) - tot(pts[("R", a, -1)])) / (2 * h)
    return grad.reshape(-1, 3), [pts[("R", a, +1)][1] for a in range(n3)], [pts[("R", a, -1)][1] for a in range(n3)]


def compute_Magnetic_Field_Gradient(parameters, basis0, C0, h):          # fin_diff.py:451-510
    pts = fd_points(parameters, basis0, C0, 0.0, h)
    tot = lambda p: p[0][0] + p[0][1] + p[0][2]
    grad = np.zeros(3)
    for b in range(3):
        grad[b] = np.real(tot(pts[("B", b, +1)]) - tot(pts[("B", b, -1)])) / (2 * h)
    return grad, [pts[("B", b, +1)][1] for b in range(3)], [pts[("B", b, -1)][1] for b in range(3)]
