"""CPU: the product's `infinity_emb` stand-in (modal-examples_b200/infinity_emb; reference call sites
06_gpu_and_ml/embeddings/image_embeddings_infinity.py:126-127, 296-306, 340-356) -- API surface, CLIP preprocessing against
HF's CLIPImageProcessor, and that there is no CPU path (the engine fails loudly without the CUDA library / a GPU)."""
import asyncio

import numpy as np
import pytest


def test_preprocessing_matches_hf_clip_image_processor():
    from infinity_emb.preprocess import CLIP_MEAN, CLIP_STD, clip_preprocess
    from PIL import Image
    from transformers import CLIPImageProcessor

    rng = np.random.default_rng(0)
    imgs = [Image.fromarray(rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)),
            Image.fromarray(rng.integers(0, 256, (224, 224), dtype=np.uint8)),      # greyscale -> RGB
            Image.fromarray(rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)),   # landscape: resize + centre crop
            Image.fromarray(rng.integers(0, 256, (512, 260, 3), dtype=np.uint8))]   # portrait
    mine = clip_preprocess(imgs)
    proc = CLIPImageProcessor(size={"shortest_edge": 224}, crop_size={"height": 224, "width": 224}, resample=3,
                              image_mean=CLIP_MEAN.tolist(), image_std=CLIP_STD.tolist())
    ref = proc(images=imgs, return_tensors="np")["pixel_values"]
    assert mine.shape == ref.shape == (4, 3, 224, 224) and mine.dtype == np.float32
    assert np.abs(mine[:2] - ref[:2]).max() < 1e-6          # no resampling involved: identical up to fp32 rounding
    level = float((1 / 255) / CLIP_STD.min())                # one grey level after normalisation
    d = np.abs(mine[2:] - ref[2:])
    assert d.max() <= 1.01 * level and d.mean() < 0.1 * level  # bicubic implementations round differently by <= 1 level
    # uint8 arrays (HWC and CHW) are accepted like PIL images
    a = np.asarray(imgs[0])
    assert np.array_equal(clip_preprocess([a]), mine[:1]) and np.array_equal(clip_preprocess([a.transpose(2, 0, 1)]), mine[:1])


def test_engine_api_and_no_cpu_path(monkeypatch):
    import infinity_emb
    from infinity_emb import AsyncEmbeddingEngine, AsyncEngineArray, EngineArgs
    from infinity_emb.primitives import Dtype, InferenceEngine

    args = EngineArgs(model_name_or_path="openai/clip-vit-base-patch16", batch_size=100, model_warmup=False,
                      engine=InferenceEngine.torch, dtype=Dtype.float16, device="cuda")  # the reference's arguments (:298-305)
    eng = AsyncEmbeddingEngine.from_args(args)
    assert eng.engine_args is args and not eng.running and "image_embed" in eng.capabilities
    assert list(AsyncEngineArray.from_args([args]))[0].engine_args is args
    assert asyncio.run(eng.image_embed(images=[])) == ([], 0)
    with pytest.raises(RuntimeError):
        asyncio.run(eng.image_embed(images=[np.zeros((224, 224, 3), np.uint8)]))  # not started
    with pytest.raises(ValueError):
        asyncio.run(AsyncEmbeddingEngine.from_args(EngineArgs(device="cpu")).astart())
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(Exception) as e:  # b200rt.init fails: no silent fallback
            asyncio.run(eng.astart())
        assert not eng.running and infinity_emb._runtime_refs in (0, 1)
        assert "b200rt" in type(e.value).__module__ or "cuda" in str(e.value).lower() or "libb200rt" in str(e.value)
