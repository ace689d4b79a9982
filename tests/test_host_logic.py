"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from apyib_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "apyib_b200.h")).read()
    declared = set(re.findall(r"\b(apyib_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(_lib.lib, name), "libapyib_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.lib.apyib_version() == 100


def test_get_slices_bit_exact_vs_oracle():
    from oracle import apyib_oracle as orc
    from apyib_b200 import utils
    for nbf, no, nf in ((7, 5, 0), (7, 5, 1), (22, 9, 2), (86, 16, 4)):
        w = orc.Wfn(np.eye(nbf), np.arange(nbf, dtype=float), no, np.eye(nbf), np.eye(nbf), np.zeros((1,) * 4), nfzc=nf)
        for m in ("RHF", "MP2", "CID", "CISD", "MP2_SO", "CID_SO", "CISD_SO"):
            assert utils.get_slices({"method": m}, w) == tuple(orc.get_slices({"method": m}, w))
    with pytest.raises(ValueError):
        utils.get_slices({"method": "CCSD"}, w)


def test_det_enumeration_and_index_lists_bit_exact():
    from oracle import apyib_oracle as orc
    from apyib_b200._lib import lib, check
    for no, nf, nv in ((2, 0, 2), (3, 1, 3), (5, 0, 2), (9, 2, 13), (4, 0, 1), (1, 0, 3)):
        ns, nd = C.c_int64(), C.c_int64()
        check(lib.apyib_det_enumeration(no, nf, nv, None, C.byref(ns), None, C.byref(nd)))
        s = np.zeros((ns.value, 2), dtype=np.int32)
        d = np.zeros((nd.value, 4), dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        check(lib.apyib_det_enumeration(no, nf, nv, p(s), None, p(d), None))
        so, do = orc.det_index_tables(no, nf, nv)
        assert np.array_equal(s, so) and np.array_equal(d, do)
        if nd.value:
            out = np.zeros((nd.value, no), dtype=np.int32)
            check(lib.apyib_det_index_lists(no, p(d), nd.value, 2, p(out)))
            want = np.tile(np.arange(no, dtype=np.int32), (nd.value, 1))
            want[np.arange(nd.value), d[:, 0]] = d[:, 1] + no
            want[np.arange(nd.value), d[:, 2]] = d[:, 3] + no
            assert np.array_equal(out, want)


def test_so_index_lists_sequential_swaps():
    from oracle import apyib_oracle as orc
    from apyib_b200._lib import lib, check
    nso, nocc = 8, 4
    tuples = np.array([(i, a, j, b) for i in range(nocc) for a in range(nocc, nso) for j in range(nocc)
                       for b in range(nocc, nso)], dtype=np.int32)
    out = np.zeros((len(tuples), nocc), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    check(lib.apyib_so_index_lists(nso, nocc, p(tuples), len(tuples), 2, p(out)))
    want = np.array([orc._swap_perm(nso, t)[:nocc] for t in tuples])
    assert np.array_equal(out, want)


def test_argument_errors_are_reported_not_crashed():
    from apyib_b200._lib import lib
    rc = lib.apyib_det_outer(None, 4, 40, None, 1, None, 1, None, None)
    assert rc < 0 and b"null" in lib.apyib_last_error()
    b = (C.c_int32 * 16)()
    assert lib.apyib_get_slices(3, 5, 0, 0, b) < 0
    # the entry points added for stacks / prefix sharing validate before they launch (no device needed for that)
    one = C.c_void_p(8)                                     # non-null dummy, never dereferenced on these paths
    assert lib.apyib_det_matvec_pairs(None, 22, 9, 2, one, 10, one, one, one, 78, 78, one, 13, one, 1, one, one, None) == -1
    assert b"null" in lib.apyib_last_error()
    assert lib.apyib_det_matvec_pairs(one, 22, 9, 2, one, 10, one, one, one, 78, 77, one, 13, one, 1, one, one, None) == -1
    assert b"group_len" in lib.apyib_last_error()
    assert lib.apyib_det_matvec_pairs(one, 22, 9, 2, one, 10, one, one, one, 100, 78, one, 13, one, 1, one, one, None) == -1
    assert b"whole number of groups" in lib.apyib_last_error()
    assert lib.apyib_det_matvec_pairs(one, 22, 9, 3, one, 10, one, one, one, 78, 78, one, 13, one, 1, one, one, None) == -1
    assert lib.apyib_det_matvec_pairs(one, 22, 9, 2, one, 10, one, one, one, 78, 78, one, 13, one, 5, one, one, None) == -1
    assert lib.apyib_det_matvec_pairs_stack(one, 0, 22, 9, 2, one, 10, one, one, one, 78, 78, one, 13, one, 0, 1, one, one, None) == -1
    assert lib.apyib_det_matvec_pairs(one, 30, 14, 2, one, 10, one, one, one, 120, 120, one, 16, one, 1, one, one, None) == -3
    assert b"n <= 12" in lib.apyib_last_error()             # APYIB_ERR_UNSUPPORTED: callers fall back to the per-matrix LU
    assert lib.apyib_det_outer_stack(one, 3, 22, 9, one, 10, one, one, None, 10, one, None) == -1      # sign without index
    assert lib.apyib_det_matvec_stack(one, 70000, 22, 9, one, 10, one, None, None, 10, one, 0, 1, one, one, None) == -1
    assert lib.apyib_det_set_pairs_variant(2) == -1 and lib.apyib_det_set_pairs_variant(0) == 0
    d4 = (C.c_int64 * 4)(3, 3, 3, 3)
    p4, s4 = (C.c_int32 * 4)(0, 1, 2, 3), (C.c_int64 * 4)(0, 0, 0, 0)
    assert lib.apyib_gather4_batch(0, one, d4, 0, 0, 81, one, d4, p4, s4, 1.0, p4, s4, 0.0, None) == -1       # nb = 0
    s_bad = (C.c_int64 * 4)(1, 0, 0, 0)
    assert lib.apyib_gather4_batch(0, one, d4, 0, 2, 81, one, d4, p4, s_bad, 1.0, p4, s4, 0.0, None) == -1    # block out of range


def test_geometry_round_trip_and_units():
    from apyib_b200.hostchem import Molecule, BOHR2ANG
    g = "O 0.0 0.1 0.2\nH 1.0 0.0 -0.5\nno_com\nunits bohr\n"
    m = Molecule.from_string(g)
    assert m.natom() == 2 and m.true_atomic_number(0) == 8
    m2 = Molecule.from_string(m.create_psi4_string_from_molecule())
    assert np.array_equal(m.geometry(), m2.geometry())
    a = Molecule.from_string("H 0 0 0\nH 0 0 %r\n" % BOHR2ANG)
    assert abs(a.geometry()[1, 2] - 1.0) < 1e-14


def test_partition_is_balanced_and_deterministic():
    from apyib_b200.parallel import partition
    from apyib_b200.fin_diff import aat_points, point_cost
    pts = aat_points(10)
    assert len(pts) == 66
    costs = [point_cost(p[0]) for p in pts]
    for world in (1, 2, 4, 8):
        own = partition(pts, costs, world)
        assert own == partition(pts, costs, world)
        loads = [sum(c for c, o in zip(costs, own) if o == r) for r in range(world)]
        assert max(loads) - min(loads) <= max(costs)
        assert sorted(set(own)) == list(range(world))


def test_synthetic_provider_is_hermitian_and_smooth():
    from apyib_b200 import hostchem as hc
    prov = hc.SyntheticProvider(6, 2, 2, seed=3)
    p = {"geom": prov.geometry_string(), "basis": "synthetic", "method": "CISD", "freeze_core": False,
         "F_el": [0.0] * 3, "F_mag": [0.0, 1e-3, 0.0], "provider": prov, "DIIS": True, "max_iterations": 50,
         "e_convergence": 1e-12, "d_convergence": 1e-12}
    H = hc.Hamiltonian(p)
    h = H.T + H.V
    assert np.abs(h - h.conj().T).max() < 1e-15 and np.iscomplexobj(h)
    w = hc.hf_wfn(H)
    E, Cm = w.solve_SCF(p)
    assert abs(np.imag(E)) < 1e-12
    assert np.abs(Cm.conj().T @ H.S @ Cm - np.eye(6)).max() < 1e-12


def _tiny(method="CISD", seed=8):
    from apyib_b200 import hostchem as hc
    prov = hc.SyntheticProvider(6, 2, 1, seed=seed)
    return lambda: {"geom": prov.geometry_string(), "basis": "synthetic", "method": method, "freeze_core": False,
                    "F_el": [0.0] * 3, "F_mag": [0.0] * 3, "provider": prov, "DIIS": True, "max_iterations": 100,
                    "e_convergence": 1e-13, "d_convergence": 1e-13}


def test_energy_only_drivers_host_logic(monkeypatch):
    """Point enumeration, parameter mutation/restoration, batching and differencing of compute_APT /
    compute_Hessian / compute_*_Gradient (fin_diff.py:27-263, 376-510) with the device solves replaced by
    the oracle's (checker only): must reproduce the oracle's serial restatement of the reference loops."""
    from apyib_b200 import fin_diff as fdm, parallel
    from oracle import fd_pipeline as fp
    calls = []

    def fake_many(parameters, wfns, print_level=0):
        calls.append(len(wfns))
        return [fp._solve(parameters, w) for w in wfns]
    monkeypatch.setattr(fdm, "correlated_many", fake_many)
    import apyib_b200.energy as en          # the MO phase fix is a device GEMM in the product (utils.compute_phase)
    monkeypatch.setattr(en, "compute_phase", lambda ndocc, nbf, ub, uC, b, C, ao_overlap=None: fp.compute_phase(ub, uC, b, C))
    mk = _tiny("CID", seed=9)
    p = mk()
    fd = fdm.finite_difference(p, None, None)
    monkeypatch.setattr(fdm.finite_difference, "BATCH_BYTES", 10 * 3 * 16 * 6 ** 4)      # force several batches
    apt = fd.compute_APT(1e-3, 1e-4)
    assert calls == [10, 10, 10, 6] and p["F_el"] == [0.0] * 3 and p["geom"].split() == mk()["geom"].split()
    assert np.abs(apt - fp.compute_APT(mk(), 1e-3, 1e-4)).max() < 1e-12
    assert np.array_equal(parallel.compute_parallel_apts(mk(), 1e-3, 1e-4), apt)
    hess = fd.compute_Hessian(1e-3)
    assert np.abs(hess - fp.compute_Hessian(mk(), 1e-3)).max() < 1e-12
    E_list, T0, C, basis, wfn = fp.energy(mk())
    fd = fdm.finite_difference(p, basis, C)
    g, pC, nC, pB, nB, pT, nT = fd.compute_Nuclear_Gradient(1e-4)
    wg, wpT, wnT = fp.compute_Nuclear_Gradient(mk(), basis, C, 1e-4)
    assert g.shape == (1, 3) and np.array_equal(g, wg) and all(np.array_equal(pT[a][2], wpT[a][2]) for a in range(3))
    g, pC, nC, pB, nB, pT, nT = fd.compute_Magnetic_Field_Gradient(1e-4)
    wg, wpT, wnT = fp.compute_Magnetic_Field_Gradient(mk(), basis, C, 1e-4)
    assert g.shape == (3,) and np.array_equal(g, wg) and all(np.array_equal(nT[b][2], wnT[b][2]) for b in range(3))
    assert p["F_mag"] == [0.0] * 3 and len(pC) == len(nB) == 3


@pytest.mark.parametrize("no,nf,nv", [(9, 0, 13), (9, 2, 13), (7, 2, 6), (3, 0, 4), (4, 1, 5), (12, 1, 4), (2, 0, 3), (5, 4, 3), (6, 0, 1)])
def test_sorted_lists_have_the_group_structure_of_the_prefix_shared_lu(no, nf, nv):
    """csrc/dets_pairs.cu consumes the sorted column lists of aats.py:581-618 in groups: same n-k unsubstituted
    columns in front, then every candidate column (k = 1) / every pair c < d in lexicographic order (k = 2).
    Host-only check of apyib_det_enumeration -> apyib_det_index_lists -> apyib_det_sort_lists -> prefix_groups,
    including the parity sign and the index back into the reference's enumeration order."""
    import ctypes as C
    from apyib_b200._lib import lib, check
    from apyib_b200.aats import prefix_groups

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return a, a.ctypes.data_as(C.POINTER(C.c_int32))

    ns, nd = C.c_int64(), C.c_int64()
    check(lib.apyib_det_enumeration(no, nf, nv, None, C.byref(ns), None, C.byref(nd)))
    singles, doubles = np.zeros((ns.value, 2), dtype=np.int32), np.zeros((nd.value, 4), dtype=np.int32)
    check(lib.apyib_det_enumeration(no, nf, nv, i32(singles)[1], None, i32(doubles)[1], None))
    o = no - nf
    assert ns.value == o * nv and nd.value == (o * (o - 1) // 2) * (nv * (nv - 1) // 2)
    for k, sub in ((1, singles), (2, doubles)):
        cnt = len(sub)
        if cnt == 0:
            continue
        L = np.zeros((cnt, no), dtype=np.int32)
        check(lib.apyib_det_index_lists(no, i32(sub)[1], cnt, k, i32(L)[1]))
        srt, sign, idx = np.zeros_like(L), np.zeros(cnt), np.zeros(cnt, dtype=np.int32)
        check(lib.apyib_det_sort_lists(no, i32(L)[1], cnt, i32(srt)[1], sign.ctypes.data_as(C.POINTER(C.c_double)), i32(idx)[1]))
        assert sorted(idx.tolist()) == list(range(cnt))
        # sign = parity of the permutation that moves the substituted entries behind the others (a column permutation)
        for c in range(0, cnt, max(1, cnt // 50)):
            src = L[idx[c]].tolist()
            perm = [src.index(x) for x in srt[c].tolist()]
            inv = sum(1 for a in range(no) for b in range(a + 1, no) if perm[a] > perm[b])
            assert sign[c] == (-1.0) ** inv
        g = prefix_groups(srt, no, k)
        if no <= k:
            assert g is None
            continue
        assert g is not None
        gl, cand, nc = g
        assert nc == nv and cand.dtype == np.int32 and cand.tolist() == list(range(no, no + nv))
        assert gl == (nv if k == 1 else nv * (nv - 1) // 2) and cnt % gl == 0
        # one group per set of substituted occupied columns; frozen-core columns are never substituted
        prefixes = {tuple(srt[c, :no - k]) for c in range(cnt)}
        assert len(prefixes) == cnt // gl == (o if k == 1 else o * (o - 1) // 2)
        assert all(set(range(nf)) <= set(p) for p in prefixes)
    # a list set that is NOT of that form is rejected (the kernel is then not used)
    if nd.value > 2:
        assert prefix_groups(srt[:-1], no, 2) is None or len(srt[:-1]) % gl == 0
        bad = srt.copy()
        bad[0, 0], bad[0, 1] = bad[0, 1], bad[0, 0]
        assert prefix_groups(bad, no, 2) is None or no - 2 < 2


def test_drop_in_signatures():
    """Boundary: every class / method / function of the reference's hot-path surface exists in the
    product under the same name with the same positional parameters and defaults (extra trailing keyword
    parameters with defaults are allowed).  The snapshot is extracted from the unmodified reference with `ast`
    (tests/golden/make_signatures.py)."""
    import inspect
    import json
    import importlib
    snap = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_signatures.json")))
    assert len(snap) >= 40
    for name, ref in snap.items():
        parts = name.split(".")
        obj = importlib.import_module("apyib_b200." + parts[0])
        for part in parts[1:]:
            assert hasattr(obj, part), "missing " + name
            obj = getattr(obj, part)
        sig = inspect.signature(obj)
        params = [q for q in sig.parameters.values()]
        names = [q.name for q in params]
        nref = len(ref["args"])
        if name == "utils.compute_mo_overlap" or name == "utils.compute_phase":
            nref_cmp = nref                      # (+ optional ao_overlap=: Psi4's mixed-basis overlap is a host input)
        assert names[:nref] == ref["args"], (name, names, ref["args"])
        ndef = len(ref["defaults"])
        got_def = [q.default for q in params[nref - ndef:nref]] if ndef else []
        assert got_def == ref["defaults"], (name, got_def, ref["defaults"])
        for q in params[nref:]:
            assert q.default is not inspect.Parameter.empty or q.kind in (q.VAR_POSITIONAL, q.VAR_KEYWORD), (name, q.name)


def test_vcd_drop_in_matches_unmodified_reference():
    """apyib_b200.vcd.vcd.compute_vcd_from_input (host post-processing, BASELINE configs[2] hand-off) on the
    reference's own (H2)_2 Hessian / APT / AAT reproduces the frequencies, IR intensities and VCD rotational
    strengths the UNMODIFIED reference's vcd.py:32-136 computed from them (tests/golden/reference_vcd.json)."""
    import json
    from apyib_b200.vcd import vcd
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    lit = json.load(open(os.path.join(G, "reference_literals.json")))
    fd = json.load(open(os.path.join(G, "reference_fd_drivers.json")))
    ref = json.load(open(os.path.join(G, "reference_vcd.json")))
    assert len(ref["cases"]) == 2
    for c in ref["cases"]:
        f = [x for x in fd["cases"] if x["method"] == c["method"]][0]
        w, D, R = vcd({"geom": lit["geom"]}).compute_vcd_from_input(np.array(f["Hessian"]), np.array(f["APT"]),
                                                                    np.array(c["AAT"]), print_level=0)
        wr = np.array([np.nan if x is None else x for x in c["frequencies_cm1"]])
        assert w.shape == (6,) and np.array_equal(np.isnan(w), np.isnan(wr))
        assert np.nanmax(np.abs(w - wr)) < 1e-8
        assert np.abs(D - np.array(c["ir_intensities_kmmol"])).max() < 1e-10
        assert np.abs(R - np.array(c["rotational_strengths"])).max() < 1e-10


def test_reference_arm_and_oracle_do_not_import_the_product():
    """bench.py --impl reference times the oracle port on host inputs from the neutral `hostinputs` package: neither
    that arm nor the oracle may import apyib_b200 (the product library must not be loaded in the reference arm)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); import hostinputs; from oracle import apyib_oracle, fd_pipeline, sparse_aat, mini_psi4; "
            "print(any(m.split('.')[0] == 'apyib_b200' for m in sys.modules))" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.strip()
    assert out == "False"
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, check=True)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["product_imported"] is False and line["extrapolated"] is True
    assert line["steps"] == 1 and line["unit"] == "s/molecule" and line["higher_is_better"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert 0 < line["ms_per_step"] < 60e3 and line["value"] > 0
