#!/usr/bin/env python
"""Experiment: two replicas per GPU with split SM budgets (B200RT_INIT_SPLIT_SMS) vs one whole-GPU replica: device-resident
items/s on GPU 0, each configuration in its own process on the same box."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, threading, time
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, torch, b200rt
from oracle import bge_ref as R
b200rt.init(devices=[0], flags=(K << 16))
n_rep = b200rt.num_gpus()
g = R.BGE_BASE
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(R.make_weights(g, 0, "hf"), g))
cap = b200rt.wave_capacity_items()
n_step = 2048
reps = []
for r in range(n_rep):
    reps.append(dict(ids=torch.from_numpy(R.synth_ids(n_step, 512, seed=r)).cuda(), lens=torch.full((n_step,), 512, dtype=torch.int32, device="cuda"),
                     out=torch.empty((n_step, 768), dtype=torch.float32, device="cuda"), stream=torch.cuda.Stream()))
def loop(r, steps):
    d = reps[r]
    for _ in range(steps):
        for i in range(0, n_step, cap):
            model.embed_device(r, d["ids"][i:].data_ptr(), d["lens"][i:].data_ptr(), min(cap, n_step - i), 512, d["out"][i:].data_ptr(), d["stream"].cuda_stream)
def run(steps):
    th = [threading.Thread(target=loop, args=(r, steps)) for r in range(n_rep)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    return time.perf_counter() - t0
run(2)
steps = 8 // n_rep
dt = run(steps)
ok = all(bool(torch.allclose(torch.linalg.vector_norm(d["out"], dim=1), torch.ones(n_step, device="cuda"), atol=1e-3)) for d in reps)
print(json.dumps(dict(K=K, replicas=n_rep, items_per_s=steps * n_step * n_rep / dt, ok=ok)))
'''
for rep in range(2):
    for K in (0, 40, 32, 48, 56):
        code = f"ROOT={ROOT!r}\nK={K}\n" + CHILD
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else ("FAILED " + r.stderr[-800:]), flush=True)
