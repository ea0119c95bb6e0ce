"""CPU: the product-side weight path (b200rt.weights) -- blob layout, seeded weights and the Hugging Face checkpoint
mapping -- against the oracle's independent statement of the same layout (oracle/bge_ref.py, itself pinned to HF
BertModel by tests/test_oracle.py).  Reference: download_model / TEI weight loading,
06_gpu_and_ml/embeddings/text_embeddings_inference.py:54-56."""
import os

import numpy as np
import pytest

from oracle import bge_ref as R
from b200rt import weights as W

SMALL = dict(vocab=97, hidden=768, layers=2, heads=12, inter=3072, max_pos=40, type_vocab=2, eps=1e-12)


def _geom(d):
    return R.BertGeometry(**d)


def test_layout_and_seeded_blob_agree_with_the_oracle():
    g = _geom(SMALL)
    assert W.blob_layout(SMALL) == [(n, tuple(s)) for n, s in R.blob_layout(g)]
    assert W.blob_numel(W.BGE_BASE_GEOMETRY) == R.blob_numel(R.BGE_BASE) == 108_891_648
    for seed in (0, 5):
        assert np.array_equal(W.random_blob(SMALL, seed), R.pack_blob(R.make_weights(g, seed, "hf"), g))


def test_hf_state_dict_loader_matches_oracle_mapping():
    g = _geom(SMALL)
    flat = R.make_weights(g, 7, "trained")
    sd = R.flat_to_hf_state(flat, g)
    geo, blob = W.load_hf_state_dict(sd)
    assert geo == SMALL
    assert np.array_equal(blob, R.pack_blob(flat, g))
    # prefixed checkpoints (BertForMaskedLM / sentence-transformers) and torch tensors
    import torch

    sd2 = {"bert." + k: torch.from_numpy(np.ascontiguousarray(v)).half() for k, v in sd.items()}
    sd2["cls.predictions.bias"] = torch.zeros(3)
    geo2, blob2 = W.load_hf_state_dict(sd2)
    assert geo2 == SMALL
    assert np.allclose(blob2, blob, rtol=1e-3, atol=1e-4)


def test_real_hf_model_round_trip():
    """The loader on an actual transformers.BertModel state dict: the blob it produces, fed to the oracle, reproduces
    that model's own outputs."""
    g = _geom(SMALL)
    flat = R.make_weights(g, 2, "trained")
    hf = R.build_hf_model(flat, g)
    geo, blob = W.load_hf_state_dict(hf.state_dict())
    assert geo == SMALL
    assert np.array_equal(blob, R.pack_blob(flat, g))


def test_safetensors_round_trip_and_hub_layout(tmp_path):
    g = _geom(SMALL)
    flat = R.make_weights(g, 9, "trained")
    sd = {k: np.ascontiguousarray(v) for k, v in R.flat_to_hf_state(flat, g).items()}
    snap = tmp_path / "models--BAAI--bge-base-en-v1.5" / "snapshots" / "abc123"
    snap.mkdir(parents=True)
    W.write_safetensors(str(snap / "model.safetensors"), sd)
    (snap / "config.json").write_text('{"layer_norm_eps": 1e-12, "num_attention_heads": 12}')
    assert W.resolve_hub_snapshot(str(tmp_path), "BAAI/bge-base-en-v1.5") == str(snap)
    assert W.resolve_hub_snapshot(str(tmp_path), "BAAI/other") is None
    geo, blob = W.load_hf_dir(str(snap))
    assert geo == SMALL and np.array_equal(blob, R.pack_blob(flat, g))
    # the reader agrees with the safetensors library where that is installed
    try:
        from safetensors.numpy import load_file
    except Exception:  # noqa: BLE001
        return
    lib = load_file(str(snap / "model.safetensors"))
    mine = W.read_safetensors(str(snap / "model.safetensors"))
    assert set(lib) == set(mine) and all(np.array_equal(lib[k], mine[k]) for k in lib)


def test_loader_rejects_wrong_shapes():
    g = _geom(SMALL)
    sd = R.flat_to_hf_state(R.make_weights(g, 1, "hf"), g)
    sd["encoder.layer.1.output.dense.weight"] = sd["encoder.layer.1.output.dense.weight"][:, :-1]
    with pytest.raises(ValueError):
        W.load_hf_state_dict(sd)
    with pytest.raises(KeyError):
        W.load_hf_state_dict({"foo": np.zeros(3)})
