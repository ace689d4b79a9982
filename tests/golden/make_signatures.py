"""Extracts the call signatures of the reference's hot-path surface from the UNMODIFIED sources with
`ast` (no import, no Psi4 needed) -> tests/golden/reference_signatures.json.  Run in the build container:
    python tests/golden/make_signatures.py
tests/test_host_logic.py::test_drop_in_signatures compares the product's classes / functions with this snapshot."""
import ast
import json
import os

REF = "/root/reference/apyib"
SURFACE = {
    "mp2_wfn.py": {"mp2_wfn": ["__init__", "solve_MP2", "solve_MP2_SO"]},
    "ci_wfn.py": {"ci_wfn": ["__init__", "solve_CID", "solve_CID_SO", "solve_CISD", "solve_CISD_SO"]},
    "aats.py": {"AAT": ["__init__", "compute_SO_det", "compute_normalization", "compute_SO_I_00", "compute_SO_I_0D",
                        "compute_SO_I_D0", "compute_SO_I_DD", "compute_SO_I_0S", "compute_SO_I_S0", "compute_SO_I_SS",
                        "compute_SO_I_SD", "compute_SO_I_DS", "compute_SO_aats", "compute_all_dets", "compute_spatial_aats"]},
    "fin_diff.py": {"finite_difference": ["__init__", "compute_Hessian", "compute_APT", "compute_AAT",
                                          "compute_Nuclear_Gradient", "compute_Magnetic_Field_Gradient"]},
    "energy.py": {None: ["energy", "phase_corrected_energy"]},
    "parallel.py": {None: ["compute_parallel_aats"]},
    "utils.py": {None: ["get_slices", "compute_F_MO", "compute_ERI_MO", "compute_F_SO", "compute_ERI_SO",
                        "solve_general_DIIS", "compute_mo_overlap", "compute_so_overlap", "compute_phase"]},
}


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [ast.literal_eval(d) for d in a.defaults]
    return {"args": names, "defaults": defaults, "line": fn.lineno}


def main():
    out = {}
    for fname, spec in SURFACE.items():
        tree = ast.parse(open(os.path.join(REF, fname)).read())
        for cls, fns in spec.items():
            if cls is None:
                body, prefix = tree.body, fname[:-3] + "."
            else:
                body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
                prefix = fname[:-3] + "." + cls + "."
            for n in body:
                if isinstance(n, ast.FunctionDef) and n.name in fns:
                    out[prefix + n.name] = sig(n)
            missing = [f for f in fns if prefix + f not in out]
            assert not missing, (fname, cls, missing)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(out), "signatures")


if __name__ == "__main__":
    main()
