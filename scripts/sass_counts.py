#!/usr/bin/env python
"""Instruction-class counts per kernel of the built library (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / LDTM / STTM),
TMA / bulk copies (UTMALDG / UBLKCP), legacy tensor-core MMAs (HMMA), packed fp32 (FFMA2) ...   usage: sass_counts.py [lib.so] > profiles/..."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tepose_b200", "libtepose_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
classes = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCCP", "UTMALDG", "UBLKCP", "UTMASTG", "SYNCS", "HMMA", "FFMA2", "FMUL2", "FFMA", "LDS", "STS", "LDG", "STG",
           "LDGSTS", "BAR", "MUFU", "SHFL", "RED", "ATOM", "LDL", "STL"]
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        for c in classes:
            if op == c or op.startswith(c + "."):
                per[cur][c] += 1
                break
print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction-class counts per kernel (static SASS)")
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    print(f"{k:60s} total={c['total']:5d}  " + "  ".join(f"{n}={c[n]}" for n in classes if c[n]))
print(f"{'ALL KERNELS':60s} total={tot['total']:5d}  " + "  ".join(f"{n}={tot[n]}" for n in classes if tot[n]))
