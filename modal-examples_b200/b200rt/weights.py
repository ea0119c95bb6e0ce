"""Product-side weight handling for ``b200rt_model_load``: the flat fp32 blob layout, the mapping from a Hugging Face
``BertModel`` checkpoint (state dict or ``model.safetensors``) to it, and seeded random weights of a geometry.

This is what stands in for the reference's ``download_model`` (``06_gpu_and_ml/embeddings/text_embeddings_inference.py:54-56``,
``snapshot_download(MODEL_ID)`` into the HF hub cache) + TEI's own weight loading: a user holding the real
``BAAI/bge-base-en-v1.5`` files calls :func:`load_safetensors` (or :func:`load_hf_dir` on the snapshot directory) and
hands the blob to :class:`b200rt.EmbedModel`.

Blob order (all fp32, HF shapes, ``y = x W^T + b`` with W ``[out, in]`` row-major; Q, K, V stacked into one ``[3H, H]``):
``emb.word [V,H] | emb.pos [P,H] | emb.type [T,H] | emb.ln.g | emb.ln.b`` then per layer
``qkv.w [3H,H] | qkv.b | ao.w [H,H] | ao.b | ln1.g | ln1.b | ff1.w [I,H] | ff1.b | ff2.w [H,I] | ff2.b | ln2.g | ln2.b``.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

BGE_BASE_GEOMETRY = dict(vocab=30522, hidden=768, layers=12, heads=12, inter=3072, max_pos=512, type_vocab=2, eps=1e-12)


def blob_layout(g: dict) -> list[tuple[str, tuple[int, ...]]]:
    h, i = g["hidden"], g["inter"]
    lay = [("emb.word", (g["vocab"], h)), ("emb.pos", (g["max_pos"], h)), ("emb.type", (g["type_vocab"], h)),
           ("emb.ln.g", (h,)), ("emb.ln.b", (h,))]
    for l in range(g["layers"]):
        p = f"l{l}."
        lay += [(p + "qkv.w", (3 * h, h)), (p + "qkv.b", (3 * h,)), (p + "ao.w", (h, h)), (p + "ao.b", (h,)),
                (p + "ln1.g", (h,)), (p + "ln1.b", (h,)), (p + "ff1.w", (i, h)), (p + "ff1.b", (i,)),
                (p + "ff2.w", (h, i)), (p + "ff2.b", (h,)), (p + "ln2.g", (h,)), (p + "ln2.b", (h,))]
    return lay


def blob_numel(g: dict) -> int:
    return sum(int(np.prod(s)) for _, s in blob_layout(g))


def random_blob(g: dict = BGE_BASE_GEOMETRY, seed: int = 0) -> np.ndarray:
    """HF default init (normal sigma 0.02, zero biases, unit LayerNorm) drawn tensor by tensor, in blob order, from
    ``numpy.random.default_rng(seed)`` -- the workload weights of BASELINE.md section 3 (bit-identical to the oracle's
    ``pack_blob(make_weights(g, seed, "hf"))``, which tests/test_weights.py checks)."""
    rng = np.random.default_rng(seed)
    out = np.empty(blob_numel(g), np.float32)
    o = 0
    for name, shape in blob_layout(g):
        n = int(np.prod(shape))
        if name.endswith(".w") or (name.startswith("emb.") and not name.startswith("emb.ln")):
            out[o:o + n] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)).reshape(-1)
        elif name.endswith(".g"):
            out[o:o + n] = 1.0
        else:
            out[o:o + n] = 0.0
        o += n
    return out


def _strip_prefix(sd: dict) -> dict:
    """Accept ``BertModel`` keys (``embeddings.*``), ``BertFor*`` keys (``bert.embeddings.*``) and sentence-transformers'
    ``0.auto_model.*``."""
    for pref in ("", "bert.", "0.auto_model.", "model.", "auto_model."):
        if pref + "embeddings.word_embeddings.weight" in sd:
            return {k[len(pref):]: v for k, v in sd.items() if k.startswith(pref)}
    raise KeyError("no BERT encoder weights found (looked for [bert.]embeddings.word_embeddings.weight)")


def geometry_from_state_dict(sd: dict, eps: float = 1e-12, heads: int | None = None) -> dict:
    sd = _strip_prefix(sd)
    vocab, hidden = sd["embeddings.word_embeddings.weight"].shape
    layers = 0
    while f"encoder.layer.{layers}.attention.self.query.weight" in sd:
        layers += 1
    return dict(vocab=int(vocab), hidden=int(hidden), layers=layers, heads=int(heads or hidden // 64),
                inter=int(sd["encoder.layer.0.intermediate.dense.weight"].shape[0]),
                max_pos=int(sd["embeddings.position_embeddings.weight"].shape[0]),
                type_vocab=int(sd["embeddings.token_type_embeddings.weight"].shape[0]), eps=float(eps))


def load_hf_state_dict(sd: dict, g: dict | None = None) -> tuple[dict, np.ndarray]:
    """HF ``BertModel`` state dict (values: numpy arrays or torch tensors, any float dtype) -> (geometry, fp32 blob).
    Restates the parameter naming of HF ``modeling_bert.py`` (BertEmbeddings :72-111, BertSelfAttention :143-207,
    BertSelfOutput :287-298, BertIntermediate :330-342, BertOutput :345-356)."""

    def arr(v):
        if hasattr(v, "detach"):  # torch tensor
            v = v.detach().to("cpu").float().numpy()
        return np.ascontiguousarray(v, dtype=np.float32)

    sd = _strip_prefix(sd)
    g = dict(g) if g is not None else geometry_from_state_dict(sd)
    out = np.empty(blob_numel(g), np.float32)
    o = 0

    def put(a, shape):
        nonlocal o
        a = arr(a)
        if a.shape != tuple(shape):
            raise ValueError(f"checkpoint tensor has shape {a.shape}, geometry needs {tuple(shape)}")
        out[o:o + a.size] = a.reshape(-1)
        o += a.size

    h, i = g["hidden"], g["inter"]
    put(sd["embeddings.word_embeddings.weight"], (g["vocab"], h))
    put(sd["embeddings.position_embeddings.weight"], (g["max_pos"], h))
    put(sd["embeddings.token_type_embeddings.weight"], (g["type_vocab"], h))
    put(sd["embeddings.LayerNorm.weight"], (h,))
    put(sd["embeddings.LayerNorm.bias"], (h,))
    for l in range(g["layers"]):
        s = f"encoder.layer.{l}."
        put(np.concatenate([arr(sd[s + f"attention.self.{n}.weight"]) for n in ("query", "key", "value")], 0), (3 * h, h))
        put(np.concatenate([arr(sd[s + f"attention.self.{n}.bias"]) for n in ("query", "key", "value")], 0), (3 * h,))
        put(sd[s + "attention.output.dense.weight"], (h, h))
        put(sd[s + "attention.output.dense.bias"], (h,))
        put(sd[s + "attention.output.LayerNorm.weight"], (h,))
        put(sd[s + "attention.output.LayerNorm.bias"], (h,))
        put(sd[s + "intermediate.dense.weight"], (i, h))
        put(sd[s + "intermediate.dense.bias"], (i,))
        put(sd[s + "output.dense.weight"], (h, i))
        put(sd[s + "output.dense.bias"], (h,))
        put(sd[s + "output.LayerNorm.weight"], (h,))
        put(sd[s + "output.LayerNorm.bias"], (h,))
    assert o == out.size
    return g, out


_ST_DTYPES = {"F32": np.float32, "F16": np.float16, "F64": np.float64, "I64": np.int64, "I32": np.int32, "U8": np.uint8}


def read_safetensors(path: str) -> dict:
    """Minimal safetensors reader (8-byte little-endian header length, JSON header, raw little-endian tensors): the
    format of the ``model.safetensors`` the reference's ``snapshot_download`` fetches.  BF16 tensors are widened to fp32."""
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n))
        base = 8 + n
        out = {}
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            b0, b1 = meta["data_offsets"]
            f.seek(base + b0)
            raw = f.read(b1 - b0)
            if meta["dtype"] == "BF16":
                u16 = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
                a = u16.view(np.float32)
            else:
                a = np.frombuffer(raw, dtype=np.dtype(_ST_DTYPES[meta["dtype"]]).newbyteorder("<"))
            out[name] = a.reshape(meta["shape"])
        return out


def write_safetensors(path: str, tensors: dict) -> None:
    """Inverse of :func:`read_safetensors` for fp32/fp16 arrays (used by the round-trip test and by tools that export
    seeded weights in the format the reference's loader expects)."""
    names = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.int64): "I64"}
    header, blobs, off = {}, [], 0
    for name, a in tensors.items():
        a = np.ascontiguousarray(a)
        header[name] = {"dtype": names[a.dtype], "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        blobs.append(a.tobytes())
        off += a.nbytes
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def load_safetensors(path: str, eps: float | None = None) -> tuple[dict, np.ndarray]:
    """``model.safetensors`` of a BERT encoder -> (geometry, blob).  ``eps``/heads come from the sibling ``config.json``
    when present (``layer_norm_eps``, ``num_attention_heads``)."""
    sd = read_safetensors(path)
    cfg_path = os.path.join(os.path.dirname(os.path.abspath(path)), "config.json")
    heads = None
    if os.path.exists(cfg_path):
        cfg = json.load(open(cfg_path))
        eps = cfg.get("layer_norm_eps", eps) if eps is None else eps
        heads = cfg.get("num_attention_heads")
    g = geometry_from_state_dict(sd, eps if eps is not None else 1e-12, heads)
    return load_hf_state_dict(sd, g)


def load_hf_dir(path: str) -> tuple[dict, np.ndarray]:
    """A local HF snapshot directory (what ``--huggingface-hub-cache`` + ``--model-id`` resolve to): picks
    ``model.safetensors``, else ``pytorch_model.bin`` (needs torch)."""
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        return load_safetensors(st)
    pt = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(pt):
        import torch

        sd = torch.load(pt, map_location="cpu", weights_only=True)
        cfg_path = os.path.join(path, "config.json")
        cfg = json.load(open(cfg_path)) if os.path.exists(cfg_path) else {}
        g = geometry_from_state_dict(sd, cfg.get("layer_norm_eps", 1e-12), cfg.get("num_attention_heads"))
        return load_hf_state_dict(sd, g)
    raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")


def resolve_hub_snapshot(cache_dir: str, model_id: str) -> str | None:
    """``<cache>/models--ORG--NAME/snapshots/<rev>/`` as ``huggingface_hub.snapshot_download`` lays it out (the reference
    mounts that cache as a Volume: text_embeddings_inference.py:20-24, 29-34)."""
    root = os.path.join(cache_dir, "models--" + model_id.replace("/", "--"), "snapshots")
    if not os.path.isdir(root):
        return None
    revs = sorted(os.listdir(root))
    return os.path.join(root, revs[-1]) if revs else None


# ----------------------------------------------------------------------------------------------- CLIP ViT image tower
# Second encoder on the same scheduler: openai/clip-vit-base-patch16's vision tower as served by the
# reference's image-embedding example (06_gpu_and_ml/embeddings/image_embeddings_infinity.py:76-77, 298-306).
# Blob order (fp32): patch.w [H, 3*p*p] (conv weight flattened channel, row, column) | cls [H] | pos [T, H] | pre.g | pre.b |
# per layer: ln1.g ln1.b qkv.w [3H,H] qkv.b ao.w ao.b ln2.g ln2.b ff1.w [I,H] ff1.b ff2.w [H,I] ff2.b | post.g post.b | proj.w [P, H].

CLIP_VIT_B16_GEOMETRY = dict(image=224, patch=16, hidden=768, layers=12, heads=12, inter=3072, proj=512, eps=1e-5)


def vit_blob_layout(g: dict) -> list[tuple[str, tuple[int, ...]]]:
    h, i = g["hidden"], g["inter"]
    tokens = (g["image"] // g["patch"]) ** 2 + 1
    lay = [("patch.w", (h, 3 * g["patch"] * g["patch"])), ("cls", (h,)), ("pos", (tokens, h)), ("pre.g", (h,)), ("pre.b", (h,))]
    for l in range(g["layers"]):
        p = f"l{l}."
        lay += [(p + "ln1.g", (h,)), (p + "ln1.b", (h,)), (p + "qkv.w", (3 * h, h)), (p + "qkv.b", (3 * h,)),
                (p + "ao.w", (h, h)), (p + "ao.b", (h,)), (p + "ln2.g", (h,)), (p + "ln2.b", (h,)),
                (p + "ff1.w", (i, h)), (p + "ff1.b", (i,)), (p + "ff2.w", (h, i)), (p + "ff2.b", (h,))]
    lay += [("post.g", (h,)), ("post.b", (h,)), ("proj.w", (g["proj"], h))]
    return lay


def vit_blob_numel(g: dict) -> int:
    return sum(int(np.prod(s)) for _, s in vit_blob_layout(g))


def random_vit_blob(g: dict = CLIP_VIT_B16_GEOMETRY, seed: int = 0) -> np.ndarray:
    """Seeded weights of the geometry: normal sigma 0.02 matrices / embeddings, zero biases, unit LayerNorm."""
    rng = np.random.default_rng(seed)
    out = np.empty(vit_blob_numel(g), np.float32)
    o = 0
    for name, shape in vit_blob_layout(g):
        n = int(np.prod(shape))
        if name.endswith(".g"):
            out[o:o + n] = 1.0
        elif name.endswith(".b"):
            out[o:o + n] = 0.0
        else:
            out[o:o + n] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)).reshape(-1)
        o += n
    return out


def load_clip_vision_state_dict(sd: dict, eps: float = 1e-5, heads: int | None = None) -> tuple[dict, np.ndarray]:
    """HF ``CLIPModel`` / ``CLIPVisionModelWithProjection`` state dict -> (geometry, fp32 blob).  Parameter names as in HF
    ``modeling_clip.py`` (``vision_model.embeddings.*``, ``vision_model.encoder.layers.N.*``, ``visual_projection.weight``)."""

    def arr(v):
        if hasattr(v, "detach"):
            v = v.detach().to("cpu").float().numpy()
        return np.ascontiguousarray(v, dtype=np.float32)

    pw = arr(sd["vision_model.embeddings.patch_embedding.weight"])
    h, _, p, _ = pw.shape
    tokens = sd["vision_model.embeddings.position_embedding.weight"].shape[0]
    layers = 0
    while f"vision_model.encoder.layers.{layers}.self_attn.q_proj.weight" in sd:
        layers += 1
    grid = int(round((tokens - 1) ** 0.5))
    g = dict(image=grid * p, patch=int(p), hidden=int(h), layers=layers, heads=int(heads or h // 64),
             inter=int(sd["vision_model.encoder.layers.0.mlp.fc1.weight"].shape[0]), proj=int(sd["visual_projection.weight"].shape[0]), eps=float(eps))
    out = np.empty(vit_blob_numel(g), np.float32)
    o = 0

    def put(a, shape):
        nonlocal o
        a = arr(a)
        if a.shape != tuple(shape):
            raise ValueError(f"checkpoint tensor has shape {a.shape}, geometry needs {tuple(shape)}")
        out[o:o + a.size] = a.reshape(-1)
        o += a.size

    i = g["inter"]
    put(pw.reshape(h, 3 * p * p), (h, 3 * p * p))
    put(sd["vision_model.embeddings.class_embedding"], (h,))
    put(sd["vision_model.embeddings.position_embedding.weight"], (tokens, h))
    put(sd["vision_model.pre_layrnorm.weight"], (h,))
    put(sd["vision_model.pre_layrnorm.bias"], (h,))
    for l in range(layers):
        s = f"vision_model.encoder.layers.{l}."
        put(sd[s + "layer_norm1.weight"], (h,))
        put(sd[s + "layer_norm1.bias"], (h,))
        put(np.concatenate([arr(sd[s + f"self_attn.{n}.weight"]) for n in ("q_proj", "k_proj", "v_proj")], 0), (3 * h, h))
        put(np.concatenate([arr(sd[s + f"self_attn.{n}.bias"]) for n in ("q_proj", "k_proj", "v_proj")], 0), (3 * h,))
        put(sd[s + "self_attn.out_proj.weight"], (h, h))
        put(sd[s + "self_attn.out_proj.bias"], (h,))
        put(sd[s + "layer_norm2.weight"], (h,))
        put(sd[s + "layer_norm2.bias"], (h,))
        put(sd[s + "mlp.fc1.weight"], (i, h))
        put(sd[s + "mlp.fc1.bias"], (i,))
        put(sd[s + "mlp.fc2.weight"], (h, i))
        put(sd[s + "mlp.fc2.bias"], (h,))
    put(sd["vision_model.post_layernorm.weight"], (h,))
    put(sd["vision_model.post_layernorm.bias"], (h,))
    put(sd["visual_projection.weight"], (g["proj"], h))
    assert o == out.size
    return g, out
