#!/usr/bin/env python
"""profiles/roofline_traffic.json from an .ncu-rep of tools/ncu_target.py (read here, no GPU): DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of each kernel of ONE forward, named as bench.py names them.
usage: ncu_traffic.py rep batch_items layers > profiles/roofline_traffic.json"""
import csv, io, json, subprocess, sys

rep, items, layers = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
idx = {h: i for i, h in enumerate(rows[0])}
units = rows[1]


def to_bytes(r, name):
    v, u = float(r[idx[name]]), units[idx[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


def role(r):
    n = r[idx["Kernel Name"]]
    for key, name in (("gemm_pair_kernel<0>", "gemm_qkv"), ("gemm_pair_kernel<1>", "gemm_ffn1_gelu"), ("gemm_pair_kernel<2>", "gemm_res"),
                      ("attention_kernel", "attention"), ("embed_kernel", "embed"), ("pool_normalize", "pool_normalize"),
                      ("scatter", "scatter")):
        if key in n:
            return name
    return None


out = {"_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from {rep} (ncu --set full, one wave of {items} x 512 tokens, "
                f"{layers}-layer model); durations are under ncu (cold cache, serialised)", "_batch_items": items, "_detail": {}}
seen = {}
for r in rows[2:]:
    k = role(r)
    if k:
        seen.setdefault(k, []).append(r)  # the last launch of each role is reported
if "gemm_res" in seen:  # the residual epilogue serves attn-out (K = 768) and FFN2 (K = 3072): tell them apart by duration
    rs = sorted(seen.pop("gemm_res"), key=lambda r: float(r[idx["gpu__time_duration.sum"]]))
    seen["gemm_attn_out"], seen["gemm_ffn2"] = [rs[0]], [rs[-1]]
for name, rs in seen.items():
    r = rs[-1]
    b = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    out[name] = int(b)
    out["_detail"][name] = {"dram_bytes_per_launch": int(b), "ncu_duration_us": float(r[idx["gpu__time_duration.sum"]]),
                            "tensor_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])}
print(json.dumps(out, indent=1))
