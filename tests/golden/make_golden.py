#!/usr/bin/env python
"""Generates tests/golden/bge_golden.npz by running the Python reference library -- Hugging Face
``transformers.BertModel`` (the encoder underneath the reference's in-tree torch path for a BGE model,
``06_gpu_and_ml/gpu_snapshot.py:52-59``) -- on seeded weights and inputs, fp32 on CPU, followed by CLS
pooling and L2 normalisation.  The reference itself pins no embedding value, so these
vectors pin the oracle's numpy restatement and, through it, the CUDA path.

    python tests/golden/make_golden.py          # rewrites bge_golden.npz (needs torch + transformers)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bge_ref as R  # noqa: E402

CASES = {
    # name: (layers, style, weight seed, input spec)
    "A": (12, "hf", 0, ("full", 4, 64, 0)),
    "B": (12, "hf", 0, ("ragged", 5, 96, 1, 1)),
    "C": (2, "trained", 3, ("ragged", 3, 200, 5, 3)),
    "D": (12, "hf", 0, ("full", 2, 512, 0)),
    "E": (12, "trained", 0, ("ragged", 4, 512, 1, 16)),
    # trained-like statistics (outlier channels, peaked attention, large LayerNorm gains): see oracle make_weights("hard")
    "F": (12, "hard", 0, ("ragged", 4, 512, 2, 64)),
    "G": (2, "hard", 1, ("ragged", 3, 200, 4, 8)),
}


def case_inputs(spec):
    if spec[0] == "full":
        _, n, s, seed = spec
        return R.synth_ids(n, s, seed), None
    _, n, s, seed, min_len = spec
    return R.synth_ragged(n, s, seed=seed, min_len=min_len)


def blob_digest(flat, g):
    return hashlib.sha256(R.pack_blob(flat, g).tobytes()).hexdigest()


def main():
    out = {}
    cache = {}
    for name, (layers, style, wseed, spec) in CASES.items():
        g = R.BertGeometry(layers=layers)
        key = (layers, style, wseed)
        if key not in cache:
            flat = R.make_weights(g, wseed, style)
            cache[key] = (flat, R.build_hf_model(flat, g), blob_digest(flat, g))
        flat, model, digest = cache[key]
        ids, lens = case_inputs(spec)
        emb = R.forward_hf(model, ids, lens)
        out[f"{name}_emb"] = emb
        out[f"{name}_digest"] = np.frombuffer(bytes.fromhex(digest), np.uint8)
        out[f"{name}_ids_sum"] = np.array([int(ids.astype(np.int64).sum())])
        print(name, emb.shape, digest[:16], float(np.linalg.norm(emb, axis=1).mean()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bge_golden.npz"), **out)


if __name__ == "__main__":
    main()
