"""Stub-import harness for the UNMODIFIED reference (/root/reference/apyib).

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the GPU box
has no /root/reference.  It is used by tests/golden/make_golden.py to produce
the committed fixtures that pin oracle/apyib_oracle.py, and by the optional
`reference`-marked CPU tests.

psi4 and opt_einsum are not installed here; the hot-path modules only need
`opt_einsum.contract` (replaced by numpy.einsum(optimize=True)) and, for the
real-molecule runs, a tiny psi4 look-alike (oracle/mini_psi4.py, s-type Gaussians only).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "apyib"))


def load(with_mini_psi4=False):
    """Return namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    if with_mini_psi4:
        from oracle import mini_psi4
        psi4 = mini_psi4.as_module()
    else:
        psi4 = types.ModuleType("psi4")
        psi4.core = types.ModuleType("psi4.core")
    sys.modules["psi4"] = psi4
    sys.modules["psi4.core"] = psi4.core
    oe = types.ModuleType("opt_einsum")
    oe.contract = lambda *a, **k: np.einsum(*a, optimize=True, **k)
    sys.modules["opt_einsum"] = oe
    for k in [k for k in sys.modules if k == "apyib" or k.startswith("apyib.")]:
        del sys.modules[k]
    pkg = types.ModuleType("apyib")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "apyib")]
    sys.modules["apyib"] = pkg
    ns = types.SimpleNamespace()
    names = ["utils", "mp2_wfn", "ci_wfn", "aats"]
    if with_mini_psi4:
        names += ["hamiltonian", "hf_wfn", "energy", "fin_diff", "parallel", "vcd"]
    for m in names:
        setattr(ns, m, importlib.import_module("apyib." + m))
    return ns


def make_ref_aat(ns, A):
    """Build the reference's AAT object from an oracle.AATInputs by attribute
    injection (bypasses aats.AAT.__init__, which needs Psi4 for AO overlaps)."""
    obj = ns.aats.AAT.__new__(ns.aats.AAT)
    obj.parameters = {"method": A.method}
    obj.nbf, obj.ndocc, obj.nfzc = A.nbf, A.ndocc, A.nfzc
    obj.nuc_pert_strength, obj.mag_pert_strength = A.nuc_pert_strength, A.mag_pert_strength
    for k in ("overlap_uu", "overlap_up", "overlap_un", "overlap_pu", "overlap_nu", "overlap_pp",
              "overlap_pn", "overlap_np", "overlap_nn", "unperturbed_T", "nuc_pos_T", "nuc_neg_T",
              "mag_pos_T", "mag_neg_T"):
        setattr(obj, k, getattr(A, k))
    return obj
