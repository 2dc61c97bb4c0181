#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, HMMA = legacy mma.sync, FFMA2/FADD2/FMUL2 = packed
fp32x2, LDGSTS = cp.async.  Reads the in-tree library with cuobjdump (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rgb_no_more_b200", "librgbnm_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "FFMA2", "FADD2", "FMUL2", "FFMA", "LDGSTS",
        "SYNCS", "RED", "MUFU"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn, counts, total = None, collections.OrderedDict(), {}
for ln in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        total[fn] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("UTCHMMA", "UTCQMMA") and op.startswith(k)):
                counts[fn][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (static instruction counts)")
print(f"# columns: total | " + " ".join(KEYS))
for (fn, c), name in zip(counts.items(), demangle):
    name = re.sub(r"\(.*", "", name)[:110]
    print(f"{name:110s} {total[fn]:6d} | " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
