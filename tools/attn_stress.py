#!/usr/bin/env python
"""Randomised stress of the persistent attention kernel against numpy: many (B, S, lens, score scale) draws -- several
units per CTA, split and unsplit launches, sub-block counts from 1 to 8, peaked scores that force accumulator rescales --
each run twice (bitwise determinism) and checked per (item, head)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np
import b200rt

b200rt.init(1)
rng = np.random.default_rng(int(os.environ.get("STRESS_SEED", "0")))
n_cases = int(os.environ.get("STRESS_CASES", "40"))
bad_total = 0
t0 = time.time()
for case in range(n_cases):
    S = int(rng.choice([1, 7, 64, 65, 128, 129, 197, 256, 300, 384, 449, 512]))
    B = int(rng.integers(1, 70)) if S > 256 else int(rng.integers(1, 140))
    mode = rng.integers(0, 3)
    lens = np.full(B, S) if mode == 0 else (rng.integers(1, S + 1, B) if mode == 1 else np.where(rng.random(B) < 0.5, S, rng.integers(1, S + 1, B)))
    scale = float(rng.choice([0.5, 1.0, 2.0, 3.0]))
    qkv = (rng.standard_normal((B * S, 2304)) * scale).astype(np.float16)
    if rng.random() < 0.5:  # a few dominant keys late in the sequence: the running maximum jumps, the accumulator is rescaled
        for b in range(B):
            k = int(rng.integers(0, lens[b]))
            qkv[b * S + k, 768:1536] *= 6.0
    ctx, _ = b200rt.debug_attention(qkv, lens.astype(np.int32), B, S)
    ctx2, _ = b200rt.debug_attention(qkv, lens.astype(np.int32), B, S)
    q = qkv.astype(np.float32).reshape(B, S, 3, 12, 64)
    qq, kk, vv = (q[:, :, j].transpose(0, 2, 1, 3) for j in range(3))
    s = (qq @ kk.transpose(0, 1, 3, 2)).astype(np.float64) * 0.125
    s = np.where((np.arange(S)[None, :] >= lens[:, None])[:, None, None, :], -np.inf, s)
    e = np.exp(s - s.max(-1, keepdims=True))
    ref = ((e / e.sum(-1, keepdims=True)) @ vv.astype(np.float64)).transpose(0, 2, 1, 3).reshape(B, S, 768)
    got = ctx.astype(np.float64).reshape(B, S, 768)
    valid = (np.arange(S)[None, :] < lens[:, None])[:, :, None]
    err = np.abs(got - ref)
    bad = (err > 6e-3 * np.abs(ref) + 6e-3) & valid
    same = np.array_equal(ctx, ctx2)
    finite = np.isfinite(got).all()
    print(f"case {case:3d}: B={B:3d} S={S:3d} lens mode {mode} scale {scale} units {B*12} -> max err {err[np.broadcast_to(valid, err.shape)].max():.2e} bad {int(bad.sum())} deterministic {same} finite {finite}", flush=True)
    bad_total += int(bad.sum()) + (0 if same else 1) + (0 if finite else 1)
print(f"{n_cases} cases in {time.time() - t0:.0f} s, failures: {bad_total}")
b200rt.shutdown()
sys.exit(1 if bad_total else 0)
