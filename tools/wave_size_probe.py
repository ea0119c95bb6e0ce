#!/usr/bin/env python
"""Diagnostics: device time per item of one wave as a function of the wave size.  With 148 SMs (74 CTA pairs) a wave of B
512-token items is 6B / 18B / 24B GEMM tiles and 12B attention units: B = 148 makes every kernel a whole number of rounds
(12 / 36 / 48 per pair, 12 per SM); B = 128 leaves 10.4 -> 11 rounds in attn-out / FFN2 / attention."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, b200rt
from oracle import bge_ref as R
b200rt.init(1, wave_items=WAVE)
g = R.BGE_BASE
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(R.make_weights(g, 0, "hf"), g))
prof = model.profile_forward(WAVE, 512, iters=40)
print(json.dumps(dict(wave=WAVE, prof=prof, total=sum(prof.values()), us_per_item=1e3 * sum(prof.values()) / WAVE)))
'''
for rep in range(2):
    for wave in (128, 148, 111, 74):
        code = f"ROOT={ROOT!r}\nWAVE={wave}\n" + CHILD
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        if r.returncode != 0:
            print(wave, "FAILED", r.stderr[-800:]); continue
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"wave {wave}: {d['total']:.3f} ms = {d['us_per_item']:.2f} us/item -> {1e6 / d['us_per_item']:.0f} items/s", {k: round(v, 3) for k, v in d["prof"].items()}, flush=True)
