#!/usr/bin/env python
"""Diagnostics: attention kernel vs fp64 numpy, error located per (item, head, query tile)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np
import b200rt

b200rt.init(1)
for B, S, lens, scale in [(4, 512, [512, 1, 128, 385], 2.0), (13, 512, [512] * 13, 1.0), (14, 512, [512, 100, 300] * 4 + [64, 65], 1.0), (40, 128, list(range(1, 41)), 1.0)]:
    rng = np.random.default_rng(B * 1000 + S)
    qkv = (rng.standard_normal((B * S, 2304)) * scale).astype(np.float16)
    ctx, ms = b200rt.debug_attention(qkv, np.array(lens, np.int32), B, S)
    q = qkv.astype(np.float64).reshape(B, S, 3, 12, 64)
    qq, kk, vv = (q[:, :, j].transpose(0, 2, 1, 3) for j in range(3))
    s = (qq @ kk.transpose(0, 1, 3, 2)) * 0.125
    s = np.where((np.arange(S)[None, :] >= np.array(lens)[:, None])[:, None, None, :], -np.inf, s)
    e = np.exp(s - s.max(-1, keepdims=True))
    ref = ((e / e.sum(-1, keepdims=True)) @ vv).transpose(0, 2, 1, 3).reshape(B, S, 12, 64)
    got = ctx.astype(np.float64).reshape(B, S, 12, 64)
    bad = np.abs(got - ref) > 4e-3 * np.abs(ref) + 4e-3
    nq = (S + 127) // 128
    print(f"B={B} S={S} lens={lens[:6]}... units={B*12} (split={B*12<148}) bad={int(bad.sum())} max_err={np.abs(got-ref).max():.3e}")
    for b in range(B):
        for h in range(12):
            for t in range(nq):
                blk = bad[b, t * 128:(t + 1) * 128, h]
                if blk.any():
                    rows = np.where(blk.any(1))[0]
                    u = (b * 12 + h) * nq + t if B * 12 < 148 else b * 12 + h
                    print(f"   item {b} head {h} tile {t} unit {u} (cta {u % 148}, seq {u // 148}): {int(blk.sum())} bad, rows {rows.min()}..{rows.max()} ({len(rows)}), max err {np.abs(got-ref)[b, t*128:(t+1)*128, h].max():.3f}")
b200rt.shutdown()
