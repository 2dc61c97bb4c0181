#!/usr/bin/env python
"""Per-step kernel breakdown from an `ncu --metrics gpu__time_duration.sum` launch list of bench.py --no-graph.
usage: tools/launch_breakdown.py gpurun_out/launches.csv [out.txt]   (one step = from one k0_dcstats launch to the next)"""
import collections
import csv
import sys

rows, h = [], None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        h = r
        continue
    if h and len(r) == len(h):
        rows.append(dict(zip(h, r)))
marks = [i for i, d in enumerate(rows) if "k0_dcstats" in d["Kernel Name"]]
start, end = marks[-2], marks[-1]          # the last complete step
sel = rows[start:end]
tot = sum(float(d["Metric Value"]) for d in sel)
agg = collections.defaultdict(list)
for d in sel:
    agg[(d["Kernel Name"][:100], d["Grid Size"])].append(float(d["Metric Value"]))
lines = [f"# {sys.argv[1]}: last complete eager step = {len(sel)} launches, {tot / 1e3:.1f} us of kernel time "
         "(ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised -- compare shares)"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{sum(v) / tot * 100:5.1f}%  n={len(v):3d} avg={sum(v) / len(v) / 1e3:8.2f} us  tot={sum(v) / 1e3:8.1f} us  {k[0]} grid={k[1]}")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
