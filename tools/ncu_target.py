#!/usr/bin/env python
"""Small ncu target: load the BGE-base-geometry model and run a few device-resident forwards of one
wave (64 x 512 tokens).  Used as `ncu ... python tools/ncu_target.py` (never timed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np  # noqa: E402
import b200rt  # noqa: E402
from oracle import bge_ref as R  # noqa: E402

layers = int(os.environ.get("NCU_LAYERS", "2"))
items = int(os.environ.get("NCU_ITEMS", "64"))
b200rt.init(devices=[0])
g = R.BertGeometry(layers=layers)
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(R.make_weights(g, 0, "hf"), g))
ids = R.synth_ids(items, 512, 0)
for _ in range(int(os.environ.get("NCU_ITERS", "2"))):
    out = model.embed(ids)
print("ok", float(np.linalg.norm(out, axis=1).mean()))
b200rt.shutdown()
