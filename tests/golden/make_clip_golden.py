#!/usr/bin/env python
"""Generates tests/golden/clip_golden.npz by running the Python reference library -- Hugging Face
``transformers.CLIPVisionModelWithProjection`` (the tower behind the reference's image-embedding example,
``06_gpu_and_ml/embeddings/image_embeddings_infinity.py:76-77``: ``openai/clip-vit-base-patch16``) -- on seeded weights and
seeded pixel inputs, fp32 on CPU, followed by L2 normalisation.  The reference pins no embedding value, so these vectors pin
the oracle's numpy restatement (oracle/clip_ref.py) and, through it, the CUDA path.

    python tests/golden/make_clip_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import clip_ref as C  # noqa: E402

CASES = {
    # name: (layers, style, weight seed, n images, pixel seed)
    "VA": (2, "trained", 1, 3, 2),
    "VB": (12, "hf", 0, 4, 0),
    "VC": (12, "trained", 2, 5, 3),
}


def main():
    out = {}
    for name, (layers, style, wseed, n, pseed) in CASES.items():
        g = C.VitGeometry(layers=layers)
        flat = C.make_weights(g, wseed, style)
        px = C.synth_pixels(n, g, pseed)
        emb = C.forward_hf(C.build_hf_model(flat, g), px)
        out[f"{name}_emb"] = emb
        out[f"{name}_digest"] = np.frombuffer(hashlib.sha256(C.pack_blob(flat, g).tobytes()).digest(), np.uint8)
        out[f"{name}_px_sum"] = np.array([float(px.astype(np.float64).sum())])
        print(name, emb.shape, float(np.linalg.norm(emb, axis=1).mean()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_golden.npz"), **out)


if __name__ == "__main__":
    main()
