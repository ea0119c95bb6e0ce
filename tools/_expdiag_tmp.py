import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, b200rt
b200rt.LIB_PATH = os.path.join(ROOT, "modal-examples_b200", "libb200rt_expdiag.so")
b200rt.init(1)
rng = np.random.default_rng(0)
for B, S in ((13, 512), (40, 512), (128, 512)) * 6:
    qkv = rng.standard_normal((B * S, 2304)).astype(np.float16)
    ctx, ms = b200rt.debug_attention(qkv, np.full(B, S, np.int32), B, S, iters=30)
    q = qkv.astype(np.float64).reshape(B, S, 3, 12, 64)
    qq, kk, vv = (q[:, :, j].transpose(0, 2, 1, 3) for j in range(3))
    s = (qq @ kk.transpose(0, 1, 3, 2)) * 0.125
    e = np.exp(s - s.max(-1, keepdims=True))
    ref = ((e / e.sum(-1, keepdims=True)) @ vv).transpose(0, 2, 1, 3).reshape(B * S, 768)
    print(B, S, "ms", ms, "max err", np.abs(ctx.astype(np.float64) - ref).max(), flush=True)
