"""Instruction mix + stall summary of one kernel from an .ncu-rep (ncu --page source --csv).
usage: python tools/ncu_mix.py REP WARP_UNITS   (WARP_UNITS = work items per warp-instruction count normaliser)"""
import csv, collections, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
ops, wf, tot = collections.Counter(), collections.Counter(), 0
stalls = collections.Counter()
kinds = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    parts = op.split(".")
    op = parts[0] + ("." + parts[1] if parts[0] in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "IMAD") and len(parts) > 1 else "")
    n = int(r[iE]); ops[op] += n; tot += n
    try: wf[op] += int(r[iW])
    except ValueError: pass
    for k in kinds:
        stalls[k] += int(r[hdr.index(k)] or 0)
print("SASS lines %d, warp instructions %d = %.1f per unit" % (len(data), tot, tot / units))
for op, n in ops.most_common(22):
    print("%-14s %8.1f per unit %5.1f%%   smem wavefronts/unit %.1f" % (op, n / units, 100.0 * n / tot, wf[op] / units))
print("stall samples:", ", ".join("%s %d" % (k[6:], v) for k, v in stalls.most_common(10)))
