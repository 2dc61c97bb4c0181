#!/usr/bin/env python
"""Dynamic SASS opcode histogram + stall-reason totals of the first kernel in an ncu report (read here, no GPU needed).
usage: tools/ncu_ops.py report.ncu-rep units   (units = work items to normalise by, e.g. 12544 quads) [listing.txt]"""
import collections, csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
i_src, i_ex, i_s = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ops, samp, forms, st, tot, totS, out = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), 0, 0, []
for r in rows[hi + 1:]:
    if len(r) <= i_ex or not r[0].startswith("0x"):
        continue
    src, ex, s = r[i_src].strip(), int(r[i_ex] or 0), int(r[i_s] or 0)
    toks = src.split()
    full = toks[1] if toks[0].startswith("@") else toks[0]
    op = full.split(".")[0]
    ops[op] += ex; samp[op] += s; tot += ex; totS += s
    forms["IMAD.MOV" if full.startswith("IMAD.MOV") else full] += ex
    for c in stall_cols:
        v = r[h.index(c)]
        if v:
            st[c] += int(v)
    out.append(f"{len(out):5d} {ex:8d} {s:4d}  {src}")
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(out))
print(f"executed {tot} = {tot / units:.1f}/unit, stall samples {totS}")
for op, c in ops.most_common(24):
    print(f"  {op:10s} {c / units:8.1f}/unit  samples {100 * samp[op] / max(1, totS):5.1f}%")
print("  forms:", ", ".join(f"{k} {v / units:.1f}" for k, v in forms.most_common(14) if k.split('.')[0] in ("IMAD", "MOV", "LOP3", "IADD3", "SHF", "VIADD", "SEL", "BRA", "ISETP", "LEA")))
print("  stalls:", {k: v for k, v in st.most_common(10)})
