#!/usr/bin/env python
"""A/B on one box: the product library vs the experiment build (`make -C modal-examples_b200/csrc exp`) on the attention kernel
(128 x 512-token items, and a ragged batch), each in its own process; alternating runs, device time per launch."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, b200rt
b200rt.LIB_PATH = LIB
b200rt.init(1)
rng = np.random.default_rng(0)
out = {}
for name, B, S, lens in (("full", 128, 512, None), ("ragged", 128, 512, rng.integers(16, 513, 128))):
    qkv = rng.standard_normal((B * S, 2304)).astype(np.float16)
    l = np.full(B, S, np.int32) if lens is None else lens.astype(np.int32)
    ctx, ms = b200rt.debug_attention(qkv, l, B, S, iters=30)
    out[name] = ms * 1e3
    out[name + "_sum"] = float(np.abs(ctx.astype(np.float64)).sum())
print(json.dumps(out))
'''
res = {"product": [], "exp": []}
for rep in range(3):
    for key, lib in [("product", "libb200rt.so"), ("exp", "libb200rt_exp.so")] + [(os.path.basename(p)[10:-3], os.path.basename(p)) for p in sorted(__import__("glob").glob(os.path.join(ROOT, "modal-examples_b200", "libb200rt_exp_*.so")))]:
        code = f"ROOT={ROOT!r}\nLIB={os.path.join(ROOT, 'modal-examples_b200', lib)!r}\n" + CHILD
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        if r.returncode != 0:
            print(key, "FAILED", r.stderr[-1500:])
            continue
        res.setdefault(key, []).append(json.loads(r.stdout.strip().splitlines()[-1]))
        print(key, res[key][-1], flush=True)
for k, v in res.items():
    if v:
        print(k, "full us", sorted(x["full"] for x in v), "ragged us", sorted(x["ragged"] for x in v))
