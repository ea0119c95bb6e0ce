#!/usr/bin/env python
"""Diagnostics: where an e2e step of the N-replica pool spends its time -- submit/wait of 32 waves of 32-item inputs with
pinned, lent buffers (what bench.py's `e2e` times), against the scheduler's per-wave statistics: the root's forward, the
slowest replica's forward, the largest compute-stream gap any replica had, and the host-side stages."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "modal-examples_b200"))
import numpy as np, torch, b200rt
from oracle import bge_ref as R
N = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
inp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
b200rt.init(N)
g = R.BGE_BASE
model = b200rt.EmbedModel(R.geometry_dict(g), R.pack_blob(R.make_weights(g, 0, "hf"), g))
cap = b200rt.wave_capacity_items()
n = 32 * cap * N
pin_ids = b200rt.PinnedBuffer((n, 512), np.int32); pin_ids.array[:] = R.synth_ids(cap, 512, 0)[np.arange(n) % cap]
pin_out = b200rt.PinnedBuffer((n, 768), np.float32)
def step():
    tk = [model.submit(pin_ids.array[i:i + inp], None, out=pin_out.array[i:i + inp], borrow_ids=True) for i in range(0, n, inp)]
    for t in tk: model.wait(t)
for _ in range(2): step()
keys = ("stage_us", "dispatch_us", "h2d_scatter_us", "forward_us", "forward_max_us", "gap_us", "gap_max_us", "d2h_us")
for rep in range(3):
    s0 = b200rt.stats(); t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0; s1 = b200rt.stats()
    print(f"N={N} inputs of {inp}: step {dt * 1e3:.1f} ms = {n / dt:.0f} items/s; waves {s1['waves'] - s0['waves']};",
          {k[:-3]: round((s1[k] - s0[k]) / 1e3, 2) for k in keys}, flush=True)

# ---- the same waves device-resident (one enqueueing thread per replica, ids already in each replica's HBM), with SM clocks
import threading
try:
    import pynvml
    pynvml.nvmlInit()
    handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(N)]
except Exception:  # noqa: BLE001
    handles = []
def sample_clocks(stop, out):
    while not stop.is_set():
        out.append([pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM) for h in handles] + [pynvml.nvmlDeviceGetPowerUsage(h) / 1e3 for h in handles])
        time.sleep(0.02)
def with_clocks(fn):
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, out)); th.start()
    t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
    stop.set(); th.join()
    a = np.array(out[len(out) // 4:]) if out else np.zeros((1, 2 * N))
    return dt, np.median(a, 0)
bufs = []
for r in range(N):
    with torch.cuda.device(r):
        bufs.append(dict(ids=torch.from_numpy(pin_ids.array[:32 * cap].copy()).cuda(), lens=torch.full((32 * cap,), 512, dtype=torch.int32, device="cuda"),
                         out=torch.empty((32 * cap, 768), dtype=torch.float32, device="cuda"), st=torch.cuda.Stream()))
def dev_step():
    def one(r):
        b = bufs[r]
        with torch.cuda.device(r):
            for i in range(0, 32 * cap, cap):
                model.embed_device(r, b["ids"][i:].data_ptr(), b["lens"][i:].data_ptr(), cap, 512, b["out"][i:].data_ptr(), b["st"].cuda_stream)
            b["st"].synchronize()
    ths = [threading.Thread(target=one, args=(r,)) for r in range(N)]
    [t.start() for t in ths]; [t.join() for t in ths]
dev_step()
for rep in range(2):
    dt, med = with_clocks(dev_step)
    print(f"device-resident step {dt * 1e3:.1f} ms; median SM MHz {med[:N]} W {med[N:]}", flush=True)
    dt, med = with_clocks(step)
    print(f"e2e step             {dt * 1e3:.1f} ms; median SM MHz {med[:N]} W {med[N:]}", flush=True)
