"""CPU: the CLIP ViT oracle (oracle/clip_ref.py) against the committed HF golden vectors, and the product-side checkpoint
mapping (b200rt.weights.load_clip_vision_state_dict) against the oracle's independent statement of the blob layout."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

from oracle import clip_ref as C
from b200rt import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_clip_golden", os.path.join(ROOT, "tests", "golden", "make_clip_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "clip_golden.npz"))


@pytest.mark.parametrize("case", ["VA", "VB"])
def test_numpy_restatement_matches_hf_golden(golden, case):
    layers, style, wseed, n, pseed = mg.CASES[case]
    g = C.VitGeometry(layers=layers)
    flat = C.make_weights(g, wseed, style)
    assert hashlib.sha256(C.pack_blob(flat, g).tobytes()).digest() == bytes(golden[f"{case}_digest"])
    px = C.synth_pixels(n, g, pseed)
    assert float(px.astype(np.float64).sum()) == float(golden[f"{case}_px_sum"][0])
    emb = C.forward_np(flat, px[:2], g)
    assert C.rel_l2(emb, golden[f"{case}_emb"][:2]).max() < 5e-6
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-6)


def test_hf_model_reproduces_golden(golden):
    layers, style, wseed, n, pseed = mg.CASES["VA"]
    g = C.VitGeometry(layers=layers)
    flat = C.make_weights(g, wseed, style)
    emb = C.forward_hf(C.build_hf_model(flat, g), C.synth_pixels(n, g, pseed))
    assert C.rel_l2(emb, golden["VA_emb"]).max() < 1e-6


def test_product_layout_and_checkpoint_mapping_agree_with_the_oracle():
    g = C.VitGeometry(layers=2)
    gd = C.geometry_dict(g)
    assert W.vit_blob_layout(gd) == [(n, tuple(s)) for n, s in C.blob_layout(g)]
    assert W.vit_blob_numel(W.CLIP_VIT_B16_GEOMETRY) == C.blob_numel(C.CLIP_B16)
    flat = C.make_weights(g, 4, "trained")
    hf = C.build_hf_model(flat, g)
    geo, blob = W.load_clip_vision_state_dict(hf.state_dict(), eps=g.eps)
    assert geo == gd
    assert np.array_equal(blob, C.pack_blob(flat, g))
    assert np.array_equal(W.random_vit_blob(gd, 3), C.pack_blob(C.make_weights(g, 3, "hf"), g))
