#!/usr/bin/env python
"""Diagnostics: dump CTA 0's per-sub-block clock stamps of the attention kernel (B200RT_ATTN_STAMPS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
out = os.path.join(ROOT, "gpurun_out", "attn_timeline.txt")
os.makedirs(os.path.dirname(out), exist_ok=True)
os.environ["B200RT_ATTN_STAMPS"] = out
import numpy as np, b200rt
b200rt.LIB_PATH = os.path.join(ROOT, "modal-examples_b200", "libb200rt_diag.so")  # `make -C modal-examples_b200/csrc diag`
b200rt.init(1)
rng = np.random.default_rng(0)
B, S = 64, 512
qkv = rng.standard_normal((B * S, 2304)).astype(np.float16)
ctx, ms = b200rt.debug_attention(qkv, np.full(B, S, np.int32), B, S, iters=5)
print("ms", ms)
print(open(out).read())
