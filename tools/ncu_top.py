#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): per kernel the headline metrics and the SASS lines with most stall samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum"]
names = []
for r in rows[2:]:
    nm = r[idx["Kernel Name"]]
    names.append(nm)
    if pat and pat not in nm: continue
    print("====", nm[:100], "grid", r[idx["launch__grid_size"]])
    for w in want:
        if w in idx: print(f"    {w:75s} {r[idx[w]]} {rows[1][idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None; shdr = None; data = {}
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name": kern = f"{r[1][:90]}#{len(data)}"; data[kern] = []; continue
    if r and r[0] == "Address": shdr = r; continue
    if kern and len(r) > 5: data[kern].append(r)
for k, v in data.items():
    if pat and pat not in k: continue
    si = shdr.index("# Samples"); ie = shdr.index("Instructions Executed")
    tot = sum(int(x[si]) for x in v) or 1
    print("---- stall samples:", k, "total", tot)
    for i, x in sorted(enumerate(v), key=lambda t: -int(t[1][si]))[:ntop]:
        print(f"   {int(x[si]) * 100 / tot:5.1f}%  line {i:5d} exec={x[ie]:>9}  {x[1].strip()[:100]}")
