#!/usr/bin/env python
"""GEMM diagnostics on the CTA-pair kernel: normal vs MMA-only (no TMA) vs TMA-only (no MMA)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np
import b200rt
b200rt.init(1)
rng = np.random.default_rng(0)
res = []
for M, N, K, epi in [(32768, 2304, 768, 0), (32768, 768, 3072, 2), (32768, 3072, 768, 1), (32768, 768, 768, 2)]:
    a = rng.standard_normal((M, K)).astype(np.float16)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    bias = np.zeros(N, np.float32)
    resid = np.zeros((M, N), np.float32) if epi == 2 else None
    for mode, nst in [(0, 0), (1, 0), (2, 0), (0, 2), (0, 3), (0, 4), (2, 2), (2, 3), (2, 4)]:
        out = np.empty((M, N), np.float32 if epi == 2 else np.float16)
        import ctypes
        ms = ctypes.c_float(0)
        lib = b200rt.load_library()
        rc = lib.b200rt_debug_gemm(epi | ((mode | (nst << 4)) << 8), a.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p), bias.ctypes.data_as(ctypes.c_void_p),
                                   resid.ctypes.data_as(ctypes.c_void_p) if resid is not None else None, out.ctypes.data_as(ctypes.c_void_p), M, N, K, 20, ctypes.byref(ms),
                                   None, None, None, 1e-12, None)
        assert rc == 0, lib.b200rt_last_error()
        c = dict(M=M, N=N, K=K, epi=epi, mode=["normal", "mma_only", "tma_only"][mode], stages=nst or 6, ms=ms.value, tflops_equiv=2.0 * M * N * K / ms.value / 1e9)
        res.append(c); print(c, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_diag.json"), "w"), indent=1)
