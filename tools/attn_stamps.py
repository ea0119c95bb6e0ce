import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
os.environ["B200RT_ATTN_STAMPS"] = "1"
import numpy as np, b200rt
b200rt.init(1)
rng = np.random.default_rng(0)
B, S = 64, 512
qkv = rng.standard_normal((B * S, 2304)).astype(np.float16)
ctx, ms = b200rt.debug_attention(qkv, np.full(B, S, np.int32), B, S, iters=5)
print("ms", ms)
