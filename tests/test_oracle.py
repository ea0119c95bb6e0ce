"""CPU: the oracle's numpy restatement against the committed golden vectors (HF BertModel outputs,
tests/golden/make_golden.py) and against HF BertModel run here; script-side restatements."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

from oracle import bge_ref as R

_spec = importlib.util.spec_from_file_location(
    "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

_W = {}


def weights(layers, style, seed):
    key = (layers, style, seed)
    if key not in _W:
        g = R.BertGeometry(layers=layers)
        _W[key] = (g, R.make_weights(g, seed, style))
    return _W[key]


def test_param_count_and_blob_layout():
    g = R.BGE_BASE
    # 108 891 648 parameters without the pooler
    assert R.blob_numel(g) == 108_891_648
    assert g.flops_per_item(512) == pytest.approx(96.64e9, rel=1e-3)


@pytest.mark.parametrize("case", ["A", "B", "C", "G"])
def test_numpy_restatement_matches_golden(golden, case):
    layers, style, wseed, spec = mg.CASES[case]
    g, flat = weights(layers, style, wseed)
    digest = hashlib.sha256(R.pack_blob(flat, g).tobytes()).digest()
    assert bytes(golden[f"{case}_digest"]) == digest, "seeded weights differ from the ones the fixtures were made with"
    ids, lens = mg.case_inputs(spec)
    assert int(ids.astype(np.int64).sum()) == int(golden[f"{case}_ids_sum"][0])
    emb = R.forward_np(flat, ids, lens, g, dtype=np.float64)
    rel = R.rel_l2(emb, golden[f"{case}_emb"])
    assert rel.max() < (2e-5 if style == "hard" else 5e-6), rel  # fp64 restatement vs HF fp32 (the trained-like style amplifies fp32 rounding ~4x)
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-6)


def test_numpy_restatement_matches_golden_512(golden):
    layers, style, wseed, spec = mg.CASES["D"]
    g, flat = weights(layers, style, wseed)
    ids, lens = mg.case_inputs(spec)
    emb = R.forward_np(flat, ids[:1], None, g, dtype=np.float32)
    assert R.rel_l2(emb, golden["D_emb"][:1]).max() < 2e-5


def test_hf_model_reproduces_golden(golden):
    """The fixtures are reproducible from the Python reference library installed here."""
    layers, style, wseed, spec = mg.CASES["C"]
    g, flat = weights(layers, style, wseed)
    ids, lens = mg.case_inputs(spec)
    emb = R.forward_hf(R.build_hf_model(flat, g), ids, lens)
    assert R.rel_l2(emb, golden["C_emb"]).max() < 1e-6


def test_padding_does_not_leak():
    g, flat = weights(2, "trained", 3)
    ids, lens = R.synth_ragged(3, 40, seed=11, min_len=2)
    a = R.forward_np(flat, ids, lens, g)
    ids2 = ids.copy()
    for i, n in enumerate(lens):
        ids2[i, n:] = 2000 + i  # garbage past the length
    b = R.forward_np(flat, ids2, lens, g)
    assert np.array_equal(a, b)
    # and an item embedded alone, unpadded, gives the same vector
    for i, n in enumerate(lens):
        solo = R.forward_np(flat, ids[i : i + 1, :n], None, g)
        assert R.rel_l2(solo, a[i : i + 1]).max() < 1e-6


def test_generate_batches_drops_remainder():
    # reference text_embeddings_inference.py:156-163 has no trailing yield
    data = [(i, f"t{i}") for i in range(100)]
    batches = list(R.generate_batches(data, 32))
    assert [len(b) for b in batches] == [32, 32, 32]
    assert batches[0][0] == (0, "t0") and batches[-1][-1] == (95, "t95")
    assert list(R.generate_batches([], 32)) == []


def test_synth_inputs_shape():
    ids = R.synth_ids(5, 512, 0)
    assert ids.dtype == np.int32 and ids.shape == (5, 512)
    assert (ids[:, 0] == 101).all() and (ids[:, -1] == 102).all()
    assert ids[:, 1:-1].min() >= 1000 and ids.max() < 30522
    ids, lens = R.synth_ragged(7, 33, seed=1)
    for i, n in enumerate(lens):
        assert ids[i, 0] == 101 and (ids[i, n:] == 0).all()
