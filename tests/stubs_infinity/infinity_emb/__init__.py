"""TEST INFRASTRUCTURE (not product): a CPU stand-in for the `infinity_emb` engine whose vectors come from the ORACLE (HF
CLIPVisionModelWithProjection via oracle/clip_ref.py, seeded weights) on pixels preprocessed by HF's own
CLIPImageProcessor.  It lets the reference's image_embeddings_infinity.py run verbatim through the `modal` shim on a box
without a GPU (tests/test_sibling_scripts.py); the product's adapter (modal-examples_b200/infinity_emb) has no CPU path.
Every call is appended to $FAKE_INFINITY_LOG as one JSON line {"n": ..., "sizes": [...], "head": [[4 floats] ...]}."""
import dataclasses
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)

_lock = threading.Lock()
_shared = {}


@dataclasses.dataclass
class EngineArgs:
    model_name_or_path: str = ""
    batch_size: int = 32
    model_warmup: bool = False
    engine: object = None
    dtype: object = None
    device: str = "cpu"


def _model():
    with _lock:
        if "m" not in _shared:
            from oracle import clip_ref as C
            from transformers import CLIPImageProcessor

            g = C.VitGeometry(layers=int(os.environ.get("FAKE_INFINITY_LAYERS", "12")))
            _shared["m"] = (C, C.build_hf_model(C.make_weights(g, 0, "hf"), g),
                            CLIPImageProcessor(size={"shortest_edge": 224}, crop_size={"height": 224, "width": 224}, resample=3,
                                               image_mean=[0.48145466, 0.4578275, 0.40821073], image_std=[0.26862954, 0.26130258, 0.27577711]))
        return _shared["m"]


class AsyncEmbeddingEngine:
    def __init__(self, args):
        self.args, self.running = args, False

    @classmethod
    def from_args(cls, args):
        return cls(args)

    async def astart(self):
        _model()
        self.running = True

    async def astop(self):
        self.running = False

    async def image_embed(self, *, images):
        assert self.running
        C, model, proc = _model()
        pixels = proc(images=list(images), return_tensors="np")["pixel_values"]
        with _lock:
            vecs = C.forward_hf(model, pixels)
            log = os.environ.get("FAKE_INFINITY_LOG")
            if log:
                with open(log, "a") as f:
                    f.write(json.dumps({"n": len(images), "sizes": [list(i.size) for i in images], "head": vecs[:, :4].tolist()}) + "\n")
        return list(vecs), len(images)
