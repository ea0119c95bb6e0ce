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


per_layer = ["gemm_qkv", "attention", "gemm_attn_out", "ln1", "gemm_ffn1_gelu", "gemm_ffn2", "ln2"]
order = ["embed_ln"] + per_layer * layers + ["pool_normalize"]
launches = [r for r in rows[2:] if "f32_to_f16" not in r[idx["Kernel Name"]] and "scatter" not in r[idx["Kernel Name"]]]
fwd = launches[-len(order):]  # the last forward of the capture
assert "embed_ln" in fwd[0][idx["Kernel Name"]] and "pool" in fwd[-1][idx["Kernel Name"]], [r[idx["Kernel Name"]][:40] for r in fwd]
out = {"_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from {rep} (ncu --set full, one wave of {items} x 512 tokens, "
                f"{layers}-layer model); durations are under ncu (cold cache, serialised)", "_batch_items": items, "_detail": {}}
for name, r in zip(order, fwd):
    b = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    out[name] = int(b)  # (later layers overwrite earlier ones: same shapes)
    out["_detail"][name] = {"dram_bytes_per_launch": int(b), "ncu_duration_us": float(r[idx["gpu__time_duration.sum"]]),
                            "tensor_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
                            "kernel": r[idx["Kernel Name"]][:60]}
print(json.dumps(out, indent=1))
