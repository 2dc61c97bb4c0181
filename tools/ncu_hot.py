#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel in an ncu report (read here, no GPU needed).
usage: tools/ncu_hot.py report.ncu-rep kernel-name-substring [top_n]"""
import csv, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# name: kernel-name regex, or "#N" = N-th profiled launch of the report (0-based)
sel = ["--launch-skip", name[1:], "--launch-count", "1"] if name.startswith("#") else ["--kernel-name", "regex:" + name]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) <= i_ex or r[0] == "Address" or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append((int(r[i_s] or 0), idx, r[i_src].strip(), int(r[i_ex] or 0)))
tot = sum(d[0] for d in data)
print("kernel", rows[hi - 1][1][:80] if hi else "", "total samples", tot, "instructions", len(data))
for s, idx, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  #{idx:5d} ex={ex:9d}  {src[:120]}")
