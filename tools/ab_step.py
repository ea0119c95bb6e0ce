#!/usr/bin/env python
"""A/B on one box: product library vs the experiment build on WHOLE steps (32 back-to-back 148-item forwards, device
resident, CUDA events) and on one item's latency (graph replay) -- what per-kernel timing cannot show, e.g. launch overlap."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, torch, b200rt
b200rt.LIB_PATH = LIB
from oracle import bge_ref as R
b200rt.init(1)
g = R.BGE_BASE
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(R.make_weights(g, 0, "hf"), g))
cap = b200rt.wave_capacity_items(); n = 32 * cap
ids = torch.from_numpy(R.synth_ids(cap, 512, 0)[np.arange(n) % cap]).cuda(); lens = torch.full((n,), 512, dtype=torch.int32, device="cuda")
out = torch.empty((n, 768), dtype=torch.float32, device="cuda"); st = torch.cuda.Stream()
def step():
    for i in range(0, n, cap):
        model.embed_device(0, ids[i:].data_ptr(), lens[i:].data_ptr(), cap, 512, out[i:].data_ptr(), st.cuda_stream)
with torch.cuda.stream(st):
    for _ in range(2): step()
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); [step() for _ in range(4)]; e1.record(st); st.synchronize()
ms = e0.elapsed_time(e1) / 4
pin = b200rt.PinnedBuffer((1, 512), np.int32); pin.array[:] = R.synth_ids(1, 512, 0); po = b200rt.PinnedBuffer((1, 768), np.float32)
lat = []
for i in range(600):
    t = time.perf_counter(); model.wait(model.submit(pin.array, None, out=po.array, borrow_ids=True)); lat.append((time.perf_counter() - t) * 1e3)
lat = sorted(lat[100:])
print(json.dumps(dict(ms_per_step=ms, items_per_s=n / ms * 1e3, p50_ms=lat[len(lat) // 2], chk=float(out.float().abs().sum()), chk1=float(np.abs(po.array).sum()))))
'''
for rep in range(3):
    for key, lib in (("product", "libb200rt.so"), ("exp", "libb200rt_exp.so")):
        code = f"ROOT={ROOT!r}\nLIB={os.path.join(ROOT, 'modal-examples_b200', lib)!r}\n" + CHILD
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        if r.returncode != 0:
            print(key, "FAILED", r.stderr[-1500:]); continue
        print(key, r.stdout.strip().splitlines()[-1], flush=True)
