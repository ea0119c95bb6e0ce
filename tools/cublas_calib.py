#!/usr/bin/env python
"""Calibration only (never on the product path): what does cuBLAS reach on this box for the four GEMM shapes?"""
import json, os, sys, torch
res = []
for M, N, K in [(32768, 2304, 768), (32768, 768, 768), (32768, 3072, 768), (32768, 768, 3072), (8192, 8192, 8192)]:
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16)
    for _ in range(3):
        c = a @ w.T
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        c = a @ w.T
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    r = dict(M=M, N=N, K=K, ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
    res.append(r); print(r, flush=True)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "cublas_calib.json"), "w"), indent=1)
