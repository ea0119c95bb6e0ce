#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of libb200rt.so (evidence that the shipped binary is tcgen05/TMA code):
    python tools/sass_hist.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "modal-examples_b200", "libb200rt.so")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCATOMSWS", "SYNCS", "MUFU", "HMMA", "FFMA", "F2FP", "FMNMX3", "BAR", "LDG", "STG", "LDS", "STS", "LDL", "STL")
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
print("# cuobjdump -lelf:", ", ".join(l.split()[-1] for l in elf.strip().splitlines()))
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", cur)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        op = m.group(1)
        hist[cur]["_total"] += 1
        hist[cur][op] += 1
for fn, h in hist.items():
    print(f"\n## {fn}   ({h['_total']} instructions)")
    fam = collections.Counter()
    for op, n in h.items():
        if op == "_total":
            continue
        base = op.split(".")[0]
        if base in KEY:
            fam[op if base.startswith(("UTC", "UTMA", "LDTM", "STTM", "MUFU")) else base] += n
    for op, n in sorted(fam.items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"  {op:32s} {n}")
