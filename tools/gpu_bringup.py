#!/usr/bin/env python
"""GPU bring-up driver: runs each kernel-level check in its own process (so a trap or a hang in one
stage cannot take the others down) and writes one JSON per stage under gpurun_out/bringup/.

    gpurun -- python tools/gpu_bringup.py            # all stages
    python tools/gpu_bringup.py --stage gemm         # one stage, in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
OUT = os.path.join(ROOT, "gpurun_out", "bringup")

STAGES = ["gemm", "attn", "hidden", "full", "sched", "perf"]


def ref_gemm(np, a16, w16, bias, epi, resid=None):
    from scipy.special import erf

    acc = a16.astype(np.float32) @ w16.astype(np.float32).T + bias
    if epi == 1:
        acc64 = acc.astype(np.float64)
        acc = (acc64 * 0.5 * (1.0 + erf(acc64 / np.sqrt(2.0)))).astype(np.float32)
    if epi == 2:
        acc = acc + resid
    return acc


def stage_gemm(res):
    import numpy as np
    import b200rt

    b200rt.init(1)
    rng = np.random.default_rng(0)
    cases = [(128, 256, 64, 0), (256, 256, 128, 0), (128, 256, 768, 0), (384, 768, 768, 0), (384, 768, 768, 1), (384, 768, 768, 2),
             (1000, 768, 3072, 2), (4096, 2304, 768, 0), (4096 + 37, 3072, 768, 1), (148 * 128 * 2, 768, 768, 2)]
    res["cases"] = []
    for M, N, K, epi in cases:
        a = (rng.standard_normal((M, K)) * 1.0).astype(np.float16)
        w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
        bias = rng.standard_normal(N).astype(np.float32)
        resid = rng.standard_normal((M, N)).astype(np.float32) if epi == 2 else None
        out, ms = b200rt.debug_gemm(epi, a, w, bias, resid)
        ref = ref_gemm(np, a, w, bias, epi, resid)
        err = np.abs(out.astype(np.float32) - ref)
        tol = 2e-3 * np.abs(ref) + 2e-3 if epi != 2 else 1e-4 * np.abs(ref) + 1e-4
        bad = int((err > tol).sum())
        c = dict(M=M, N=N, K=K, epi=epi, max_abs_err=float(err.max()), ref_absmax=float(np.abs(ref).max()), n_bad=bad, ok=bad == 0)
        if bad:
            idx = np.argwhere(err > tol)
            c["first_bad"] = [int(x) for x in idx[0]]
            c["bad_rows"] = sorted(set(int(x) for x in idx[:, 0]))[:16]
            c["bad_cols"] = sorted(set(int(x) for x in idx[:, 1]))[:16]
            c["sample"] = [float(out[idx[0][0], idx[0][1]]), float(ref[idx[0][0], idx[0][1]])]
        res["cases"].append(c)
        print(c, flush=True)
    res["ok"] = all(c["ok"] for c in res["cases"])


def ref_attention(np, qkv16, lens, B, S):
    qkv = qkv16.astype(np.float64).reshape(B, S, 3, 12, 64)
    q, k, v = (qkv[:, :, j].transpose(0, 2, 1, 3) for j in range(3))  # [B,12,S,64]
    s = (q @ k.transpose(0, 1, 3, 2)) * 0.125
    mask = np.arange(S)[None, :] >= np.asarray(lens)[:, None]
    s = np.where(mask[:, None, None, :], -np.inf, s)
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    p = e / e.sum(-1, keepdims=True)
    return (p @ v).transpose(0, 2, 1, 3).reshape(B * S, 768)


def stage_attn(res):
    import numpy as np
    import b200rt

    b200rt.init(1)
    rng = np.random.default_rng(1)
    cases = [(1, 128, [128]), (1, 64, [64]), (2, 512, [512, 512]), (3, 300, [300, 17, 129]), (4, 512, [512, 1, 128, 385]), (2, 7, [7, 3])]
    res["cases"] = []
    for B, S, lens in cases:
        for scale in (1.0, 3.0):
            qkv = (rng.standard_normal((B * S, 2304)) * scale).astype(np.float16)
            ctx, ms = b200rt.debug_attention(qkv, np.array(lens, np.int32), B, S)
            ref = ref_attention(np, qkv, lens, B, S)
            # only rows < S matter for every item (padded query rows are computed but arbitrary-but-finite)
            err = np.abs(ctx.astype(np.float64) - ref)
            finite = bool(np.isfinite(ctx.astype(np.float32)).all())
            tol = 4e-3 * np.abs(ref) + 4e-3
            bad = int((err > tol).sum())
            c = dict(B=B, S=S, lens=lens, scale=scale, max_abs_err=float(err.max()), finite=finite, n_bad=bad, ok=bad == 0 and finite)
            if bad:
                idx = np.argwhere(err > tol)
                c["first_bad"] = [int(x) for x in idx[0]]
                c["bad_rows"] = sorted(set(int(x) for x in idx[:, 0]))[:24]
                c["bad_cols"] = sorted(set(int(x) for x in idx[:, 1]))[:24]
                c["sample"] = [float(ctx[idx[0][0], idx[0][1]]), float(ref[idx[0][0], idx[0][1]])]
            res["cases"].append(c)
            print(c, flush=True)
    res["ok"] = all(c["ok"] for c in res["cases"])


def stage_hidden(res):
    import numpy as np
    import b200rt
    from oracle import bge_ref as R

    b200rt.init(1)
    g = R.BertGeometry(layers=2)
    flat = R.make_weights(g, 3, "trained")
    model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
    ids, lens = R.synth_ragged(3, 200, seed=5, min_len=3)
    ref_emb, hidden = R.forward_np(flat, ids, lens, g, dtype=np.float64, return_hidden=True)
    res["layers"] = []
    for L in range(0, 3):
        h = model.debug_hidden(ids, lens, L)
        errs = []
        for i, n in enumerate(lens):
            d = h[i, :n] - hidden[L][i, :n]
            errs.append(float(np.sqrt((d ** 2).sum()) / np.sqrt((hidden[L][i, :n] ** 2).sum())))
        c = dict(layer=L, rel_l2=errs, finite=bool(np.isfinite(h).all()), ok=max(errs) < 2e-3)
        res["layers"].append(c)
        print(c, flush=True)
    emb = model.embed(ids, lens)
    rel = R.rel_l2(emb, ref_emb)
    res["embed_rel_l2"] = [float(x) for x in rel]
    res["ok"] = all(c["ok"] for c in res["layers"]) and float(rel.max()) < 1e-3
    print("embed rel", rel, flush=True)


def stage_full(res):
    import numpy as np
    import torch
    import b200rt
    from oracle import bge_ref as R

    b200rt.init(1)
    g = R.BGE_BASE
    out = {}
    for style in ("hf", "trained"):
        flat = R.make_weights(g, 0, style)
        model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
        hf = R.build_hf_model(flat, g)
        torch.set_num_threads(os.cpu_count() or 8)
        ids = R.synth_ids(8, 512, 0)
        t0 = time.time()
        ref = R.forward_hf(hf, ids)
        t_ref = time.time() - t0
        emb = model.embed(ids)
        rel = R.rel_l2(emb, ref)
        ids2, lens2 = R.synth_ragged(8, 512, seed=1, min_len=16)
        ref2 = R.forward_hf(hf, ids2, lens2)
        emb2 = model.embed(ids2, lens2)
        rel2 = R.rel_l2(emb2, ref2)
        out[style] = dict(rel_l2_full=[float(x) for x in rel], rel_l2_ragged=[float(x) for x in rel2], lens=[int(x) for x in lens2],
                          cpu_ref_s=t_ref, norm=[float(x) for x in np.linalg.norm(emb, axis=1)])
        print(style, out[style], flush=True)
    res.update(out)
    res["ok"] = all(max(v["rel_l2_full"] + v["rel_l2_ragged"]) < 1e-3 for v in out.values())


def stage_sched(res):
    """Scheduler semantics: many tickets, ordered wait + unordered poll_any, split tickets, mixed lengths."""
    import numpy as np
    import b200rt
    from oracle import bge_ref as R

    n_gpus = int(os.environ.get("BRINGUP_GPUS", "1"))
    b200rt.init(n_gpus)
    g = R.BertGeometry(layers=2)
    flat = R.make_weights(g, 7, "trained")
    model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
    rng = np.random.default_rng(0)
    inputs = []
    for i in range(40):
        S = int(rng.choice([16, 64, 128, 512]))
        n = int(rng.integers(1, 48))
        ids, lens = R.synth_ragged(n, S, seed=100 + i, min_len=1)
        inputs.append((ids, lens))
    # reference = the same engine, one input at a time (bitwise determinism across batch compositions is not
    # required, only parity; compare to a one-by-one run with a tight tolerance)
    solo = [model.embed(ids, lens) for ids, lens in inputs]
    tickets = [model.submit(ids, lens, tag=i) for i, (ids, lens) in enumerate(inputs)]
    seen = {}
    for _ in range(len(tickets) // 2):
        t = model.poll_any(60000)
        assert t is not None, "poll_any timed out"
        seen[t.tag] = t.out
    for t in tickets:
        if t.tag not in seen:
            o = model.wait(t, 60000)
            assert o is not None, "wait timed out"
            seen[t.tag] = o
    worst = 0.0
    for i in range(len(inputs)):
        worst = max(worst, float(R.rel_l2(seen[i], solo[i]).max()))
    big_ids = R.synth_ids(b200rt.wave_capacity_items() * n_gpus * 2 + 5, 512, 9)  # forces a ticket to span waves
    big = model.embed(big_ids)
    chunk = model.embed(big_ids[:4])
    res.update(worst_rel_vs_solo=worst, big_norm_ok=bool(np.allclose(np.linalg.norm(big, axis=1), 1.0, atol=1e-4)),
               big_vs_chunk=float(R.rel_l2(big[:4], chunk).max()), stats=b200rt.stats(), n_gpus=n_gpus)
    # error path: id outside the vocabulary must be rejected on submit
    try:
        bad = np.full((1, 8), 40000, np.int32)
        model.submit(bad)
        res["rejects_bad_id"] = False
    except b200rt.B200RTError as e:
        res["rejects_bad_id"] = e.code == b200rt.E_INVALID
    res["ok"] = worst < 1e-4 and res["big_norm_ok"] and res["big_vs_chunk"] < 1e-4 and res["rejects_bad_id"]
    print(res, flush=True)
    b200rt.shutdown()


def stage_perf(res):
    import numpy as np
    import b200rt
    from oracle import bge_ref as R

    b200rt.init(1)
    rng = np.random.default_rng(0)
    res["gemm"] = []
    for M, N, K, epi in [(16384, 2304, 768, 0), (16384, 768, 768, 2), (16384, 3072, 768, 1), (16384, 768, 3072, 2),
                         (32768, 2304, 768, 0), (32768, 3072, 768, 1), (32768, 768, 3072, 2)]:
        a = rng.standard_normal((M, K)).astype(np.float16)
        w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
        bias = np.zeros(N, np.float32)
        resid = np.zeros((M, N), np.float32) if epi == 2 else None
        _, ms = b200rt.debug_gemm(epi, a, w, bias, resid, iters=20)
        c = dict(M=M, N=N, K=K, epi=epi, ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
        res["gemm"].append(c)
        print(c, flush=True)
    res["attn"] = []
    for B, S in [(32, 512), (64, 512), (64, 128)]:
        qkv = rng.standard_normal((B * S, 2304)).astype(np.float16)
        _, ms = b200rt.debug_attention(qkv, np.full(B, S, np.int32), B, S, iters=20)
        c = dict(B=B, S=S, ms=ms, tflops=4.0 * B * S * S * 768 / ms / 1e9)
        res["attn"].append(c)
        print(c, flush=True)
    g = R.BGE_BASE
    flat = R.make_weights(g, 0, "hf")
    model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(flat, g))
    res["profile"] = {}
    for B in (32, 64, 128):
        if B > b200rt.wave_capacity_items():
            continue
        prof = model.profile_forward(B, 512, iters=3)
        total = sum(prof.values())
        res["profile"][str(B)] = dict(per_kernel_ms=prof, total_ms=total, items_per_s=B / total * 1e3,
                                      tflops=B * g.flops_per_item(512) / total / 1e9)
        print(B, res["profile"][str(B)], flush=True)
    res["ok"] = True


def run_stage(name):
    res = {"stage": name, "ok": False}
    t0 = time.time()
    try:
        globals()["stage_" + name](res)
    except Exception as e:  # noqa: BLE001
        import traceback

        res["error"] = f"{type(e).__name__}: {e}"
        res["traceback"] = traceback.format_exc()
        print(res["traceback"], flush=True)
    res["seconds"] = time.time() - t0
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"{name}{os.environ.get('BRINGUP_TAG', '')}.json"), "w") as f:
        json.dump(res, f, indent=1)
    return res["ok"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    ap.add_argument("--stages", default=",".join(STAGES))
    ap.add_argument("--timeout", type=int, default=420)
    a = ap.parse_args()
    if a.stage:
        sys.exit(0 if run_stage(a.stage) else 1)
    os.makedirs(OUT, exist_ok=True)
    summary = {}
    for st in a.stages.split(","):
        t0 = time.time()
        with open(os.path.join(OUT, f"{st}{os.environ.get('BRINGUP_TAG', '')}.log"), "w") as log:
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], stdout=log, stderr=subprocess.STDOUT,
                                   timeout=a.timeout)
                summary[st] = {"rc": p.returncode, "s": round(time.time() - t0, 1)}
            except subprocess.TimeoutExpired:
                summary[st] = {"rc": "timeout", "s": round(time.time() - t0, 1)}
        print(st, summary[st], flush=True)
        tail = open(os.path.join(OUT, f"{st}{os.environ.get('BRINGUP_TAG', '')}.log")).read()[-2500:]
        print(tail, flush=True)
    with open(os.path.join(OUT, f"summary{os.environ.get('BRINGUP_TAG', '')}.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
