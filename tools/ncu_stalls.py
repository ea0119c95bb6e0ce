#!/usr/bin/env python
"""Per-region stall-reason breakdown of one kernel in an .ncu-rep (read here, no GPU).
usage: ncu_stalls.py rep kernel_substr [first_line last_line]  -- lines are indices into the SASS listing"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        kern = r[1]; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and pat in kern and len(r) > 5:
        rows.append(r)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
sel = rows[lo:hi + 1]
tot = sum(int(r[si]) for r in sel)
print(f"lines {lo}..{min(hi, len(rows) - 1)} of {len(rows)}: {tot} samples of {sum(int(r[si]) for r in rows)}; warp-instructions executed {sum(int(r[ie]) for r in sel)}")
for h in reasons:
    s = sum(int(r[hdr.index(h)] or 0) for r in sel)
    if s: print(f"  {h:24s} {s:7d} {100 * s / max(tot, 1):5.1f}%")
