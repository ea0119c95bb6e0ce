#!/usr/bin/env python
"""A/B on one box: product library vs experiment build on the whole forward -- per-kernel device times of 128-item waves
(12 layers, 30 back-to-back forwards) in separate processes, alternating."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, b200rt
b200rt.LIB_PATH = LIB
from oracle import bge_ref as R
b200rt.init(1)
g = R.BGE_BASE
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(R.make_weights(g, 0, "hf"), g))
prof = model.profile_forward(128, 512, iters=40)
e = model.embed(R.synth_ids(8, 512, 3))
print(json.dumps(dict(prof=prof, total=sum(prof.values()), chk=float(np.abs(e).sum()))))
'''
for rep in range(2):
    for key, lib in (("product", "libb200rt.so"), ("exp", "libb200rt_exp.so")):
        code = f"ROOT={ROOT!r}\nLIB={os.path.join(ROOT, 'modal-examples_b200', lib)!r}\n" + CHILD
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        if r.returncode != 0:
            print(key, "FAILED", r.stderr[-1500:]); continue
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(key, "total %.3f ms" % d["total"], {k: round(v, 3) for k, v in d["prof"].items()}, "chk", d["chk"], flush=True)
