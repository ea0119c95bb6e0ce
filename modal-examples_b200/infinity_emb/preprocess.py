"""CLIP image preprocessing (what ``infinity_emb``'s CLIP path gets from HF ``CLIPImageProcessor`` with the
``openai/clip-vit-base-patch16`` preprocessor_config: convert to RGB, bicubic resize of the shorter edge to 224, centre crop
224 x 224, rescale by 1/255, normalise with OpenAI's channel mean / std) -> float32 ``[n, 3, 224, 224]``.

Host-side integer/float glue in front of the GPU tower; the reference example's volume already holds 224 x 224 JPEGs
(image_embeddings_infinity.py:168-186), for which only the rescale + normalise step does any work."""
from __future__ import annotations

import numpy as np

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], np.float32)


def _to_rgb_array(img, size: int) -> np.ndarray:
    """One image (PIL.Image, or an HWC / CHW uint8 array) -> uint8 [size, size, 3]."""
    if isinstance(img, np.ndarray):
        a = img
        if a.ndim == 3 and a.shape[0] in (1, 3) and a.shape[2] not in (1, 3):
            a = a.transpose(1, 2, 0)
        if a.ndim == 2:
            a = a[:, :, None]
        if a.shape[2] == 1:
            a = np.repeat(a, 3, 2)
        if a.shape[:2] == (size, size) and a.dtype == np.uint8:
            return np.ascontiguousarray(a[:, :, :3])
        from PIL import Image

        img = Image.fromarray(np.ascontiguousarray(a[:, :, :3]).astype(np.uint8))
    if img.mode != "RGB":
        img = img.convert("RGB")
    w, h = img.size
    if (w, h) != (size, size):
        from PIL import Image

        short = min(w, h)  # shorter edge -> size, aspect kept (HF: int(size * long / short) for the other edge)
        nw, nh = (size, int(size * h / w)) if w <= h else (int(size * w / h), size)
        if short != size:
            img = img.resize((nw, nh), resample=Image.BICUBIC)
        left, top = (nw - size) // 2, (nh - size) // 2
        img = img.crop((left, top, left + size, top + size))
    return np.asarray(img, dtype=np.uint8)


def clip_preprocess(images, size: int = 224, out: np.ndarray | None = None) -> np.ndarray:
    n = len(images)
    if out is None:
        out = np.empty((n, 3, size, size), np.float32)
    scale = (1.0 / 255.0) / CLIP_STD
    shift = -CLIP_MEAN / CLIP_STD
    for i, img in enumerate(images):
        a = _to_rgb_array(img, size)
        # (x / 255 - mean) / std, channel-first
        np.multiply(a.transpose(2, 0, 1), scale[:, None, None], out=out[i], casting="unsafe")
        out[i] += shift[:, None, None]
    return out
