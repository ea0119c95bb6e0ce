"""diagnostics: where does a forward with the query-tile-split attention launch differ from the unsplit one?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'modal-examples_b200'))
import numpy as np, b200rt
from oracle import bge_ref as R
b200rt.init(devices=[0])
g = R.BertGeometry(layers=2)
flat = R.make_weights(g, 0, "hf")
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
ids = R.synth_ids(13, 512, 7)
for L in (1, 2):
    big = model.debug_hidden(ids, None, L)          # 13 items: one CTA per (item, head)
    solo = model.debug_hidden(ids[5:6], None, L)    # 1 item: split by query tile
    d = np.abs(big[5].astype(np.float64) - solo[0].astype(np.float64))
    rows = np.nonzero(d.max(1) > 0)[0]
    print("layers", L, "max abs diff", d.max(), "rows differing", len(rows), "first", rows[:12], "last", rows[-5:] if len(rows) else [])
    if len(rows):
        r = rows[0]; cols = np.nonzero(d[r] > 0)[0]
        print("  row", r, "cols differing", len(cols), cols[:16], "max", d[r].max())
    print("  per 128-row tile: rows differing", [int((d[t*128:(t+1)*128].max(1) > 0).sum()) for t in range(4)])
b200rt.shutdown()
