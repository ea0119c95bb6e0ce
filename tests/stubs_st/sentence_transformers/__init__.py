"""TEST INFRASTRUCTURE (not product): a CPU stand-in for the one `sentence_transformers` call the reference's
06_gpu_and_ml/gpu_snapshot.py makes (:41-59: `SentenceTransformer(model_id, device="cuda").encode(sentences,
normalize_embeddings=True)` == BertModel -> CLS pooling -> L2 normalise for the BGE family), backed by the ORACLE (HF
BertModel via oracle/bge_ref.py, seeded weights) on the product tokenizer's ids.  The package is not installed on this box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))


class SentenceTransformer:
    def __init__(self, model_name_or_path, device=None, **_kw):
        from oracle import bge_ref as R
        from tei_router.tokenizer import WordPiece

        self.model_id, self.device = model_name_or_path, device
        self._R = R
        self._g = R.BertGeometry(layers=int(os.environ.get("FAKE_ST_LAYERS", "12")))
        self._model = R.build_hf_model(R.make_weights(self._g, 0, "hf"), self._g)
        self._tok = WordPiece(None)

    def encode(self, sentences, normalize_embeddings=False, **_kw):
        import numpy as np

        rows = [self._tok.encode(t, self._g.max_pos, True) for t in sentences]
        lens = np.array([len(r) for r in rows], np.int32)
        ids = np.zeros((len(rows), int(lens.max())), np.int32)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = r
        assert normalize_embeddings, "the stand-in implements the normalised path only"
        return self._R.forward_hf(self._model, ids, lens)
