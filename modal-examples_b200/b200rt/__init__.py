"""ctypes binding of ``libb200rt.so`` (C ABI in ``include/b200rt.h``).

This is the host-side mirror of the reference's interface for the embed hot path: where the
reference's ``TextEmbeddingsInference.embed`` (``06_gpu_and_ml/embeddings/text_embeddings_inference.py:97-104``)
POSTs strings to a TEI server and ``.map`` (``:167``) fans batches out to cloud containers, an
:class:`EmbedModel` takes token-id batches and the library's C++ scheduler fans them out to the local
B200s.  There is no CPU path: importing works anywhere, but :func:`init` raises unless the shared
library is built and a B200 is visible.
"""
from __future__ import annotations

import ctypes
import os
import threading
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libb200rt.so")

OK, TIMEOUT = 0, 1
E_INVALID, E_STATE, E_CUDA, E_NOMEM, E_UNSUPPORTED = -1, -2, -3, -4, -5

# every symbol include/b200rt.h and include/b200rt_debug.h declare (tests check the export list)
ABI_SYMBOLS = [
    "b200rt_init", "b200rt_init_devices", "b200rt_num_gpus", "b200rt_model_load", "b200rt_submit", "b200rt_submit_ex", "b200rt_submit_pixels", "b200rt_wait",
    "b200rt_poll_any", "b200rt_embed_device", "b200rt_device_sync", "b200rt_wave_capacity_items",
    "b200rt_alloc_pinned", "b200rt_free_pinned", "b200rt_stats", "b200rt_last_error", "b200rt_shutdown",
]
DEBUG_SYMBOLS = ["b200rt_debug_gemm", "b200rt_debug_attention", "b200rt_debug_hidden", "b200rt_debug_vit_hidden", "b200rt_debug_profile_forward"]


class B200RTError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200rt error {code}: {msg}")
        self.code = code


class BertConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("vocab", "hidden", "layers", "heads", "inter", "max_pos", "type_vocab")] + [
        ("eps", ctypes.c_float)
    ]


class VitConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("image", "patch", "hidden", "layers", "heads", "inter", "proj")] + [("eps", ctypes.c_float)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("items", "waves", "tickets", "kernel_launches", "h2d_bytes", "d2h_bytes", "peer_bytes")] + [
        (n, ctypes.c_double) for n in ("stage_us", "h2d_scatter_us", "forward_us", "d2h_us", "gap_us", "dispatch_us", "forward_max_us", "gap_max_us")
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen the C ABI.  Fails loudly when the extension has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200RTError(E_STATE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(there is no CPU or PyTorch fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        u16p = ctypes.POINTER(ctypes.c_uint16)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        lib.b200rt_init.argtypes = [ctypes.c_int, ctypes.c_uint32]
        lib.b200rt_init_devices.argtypes = [i32p, ctypes.c_int, ctypes.c_uint32]
        lib.b200rt_model_load.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, i32p]
        lib.b200rt_submit.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, u64p]
        lib.b200rt_submit_ex.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_uint32, u64p]
        lib.b200rt_submit_pixels.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, u64p]
        lib.b200rt_debug_vit_hidden.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.b200rt_wait.argtypes = [ctypes.c_uint64, ctypes.c_int]
        lib.b200rt_poll_any.argtypes = [u64p, ctypes.c_int]
        lib.b200rt_embed_device.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.b200rt_device_sync.argtypes = [ctypes.c_int]
        lib.b200rt_alloc_pinned.argtypes = [ctypes.c_size_t]
        lib.b200rt_alloc_pinned.restype = ctypes.c_void_p
        lib.b200rt_free_pinned.argtypes = [ctypes.c_void_p]
        lib.b200rt_free_pinned.restype = None
        lib.b200rt_stats.argtypes = [ctypes.POINTER(Stats)]
        lib.b200rt_last_error.restype = ctypes.c_char_p
        lib.b200rt_shutdown.restype = None
        lib.b200rt_debug_gemm.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]
        lib.b200rt_debug_attention.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        lib.b200rt_debug_hidden.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.b200rt_debug_profile_forward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t,
                                                     f32p, i32p, ctypes.c_int]
        _lib = lib
        return lib


def _check(rc: int):
    if rc < 0:
        raise B200RTError(rc, (load_library().b200rt_last_error() or b"").decode(errors="replace"))
    return rc


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------- runtime

_initialised = False


def init(n_gpus: int = 1, devices=None, flags: int = 0, wave_items: int = 0) -> int:
    """Start the replica pool on ``devices`` (default ``0..n_gpus-1``).  Idempotent per process.
    ``wave_items``: 512-token items one replica takes per wave (0 = the library default: one per SM, 148 on a B200)."""
    global _initialised
    lib = load_library()
    if _initialised:
        return lib.b200rt_num_gpus()
    flags |= wave_items & 0xFFFF
    if devices is not None:
        arr = (ctypes.c_int32 * len(devices))(*devices)
        _check(lib.b200rt_init_devices(arr, len(devices), flags))
    else:
        _check(lib.b200rt_init(n_gpus, flags))
    _initialised = True
    return lib.b200rt_num_gpus()


def shutdown():
    global _initialised
    if _initialised:
        load_library().b200rt_shutdown()
        _initialised = False


def num_gpus() -> int:
    return _check(load_library().b200rt_num_gpus())


def wave_capacity_items() -> int:
    return _check(load_library().b200rt_wave_capacity_items())


def stats() -> dict:
    s = Stats()
    _check(load_library().b200rt_stats(ctypes.byref(s)))
    return s.as_dict()


class PinnedBuffer:
    """Page-locked host array (``b200rt_alloc_pinned``)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(x) for x in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = load_library().b200rt_alloc_pinned(max(nbytes, 1))
        if not self._p:
            _check(E_NOMEM)
        buf = (ctypes.c_char * nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            load_library().b200rt_free_pinned(self._p)
            self._p = None


@dataclass
class Ticket:
    id: int
    out: np.ndarray
    tag: object = None
    ids: object = None  # keeps a lent (BORROW_IDS) input alive until the ticket completes


class TicketError(B200RTError):
    """A ticket completed with an error; carries the ticket so that the caller can tell which input failed."""

    def __init__(self, code: int, msg: str, ticket: "Ticket | None"):
        super().__init__(code, msg)
        self.ticket = ticket


# b200rt_poll_any is process-global, so the registry of live tickets is too (not per model): a ticket is registered
# under the lock that also covers its b200rt_submit call, so a poll_any thread that reaps it early blocks on the lock
# until the entry exists.
_live: dict[int, Ticket] = {}
_live_lock = threading.Lock()

SUBMIT_BORROW_IDS = 1


def poll_any(timeout_ms: int = -1) -> Ticket | None:
    """Next finished ticket nobody is waiting on (any model); None on timeout.  Raises :class:`TicketError` (with the
    ticket attached) when that ticket failed."""
    lib = load_library()
    t = ctypes.c_uint64(0)
    rc = lib.b200rt_poll_any(ctypes.byref(t), timeout_ms)
    if rc == TIMEOUT:
        return None
    with _live_lock:
        tk = _live.pop(t.value, None)
    if rc < 0:
        msg = (lib.b200rt_last_error() or b"").decode(errors="replace")
        if t.value == 0:
            raise B200RTError(rc, msg)
        raise TicketError(rc, msg, tk)
    return tk


class EmbedModel:
    """A loaded BERT-geometry encoder; ``submit``/``wait``/``poll_any`` are the .map() primitives."""

    def __init__(self, geometry: dict, blob: np.ndarray):
        lib = load_library()
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self.cfg = BertConfig(geometry["vocab"], geometry["hidden"], geometry["layers"], geometry["heads"], geometry["inter"],
                              geometry["max_pos"], geometry["type_vocab"], geometry["eps"])
        self.hidden = geometry["hidden"]
        h = ctypes.c_int32(-1)
        _check(lib.b200rt_model_load(b"bert", ctypes.byref(self.cfg), _ptr(blob), blob.nbytes, ctypes.byref(h)))
        self.handle = h.value
        self._lib = lib

    def submit(self, ids: np.ndarray, lens=None, out: np.ndarray | None = None, tag=None, borrow_ids: bool = False) -> Ticket:
        """``borrow_ids``: the caller promises not to touch ``ids`` until the ticket completes; the library then takes
        no private copy and, when ``ids``/``out`` live in :class:`PinnedBuffer` memory, DMAs from/to them directly."""
        if borrow_ids and not (isinstance(ids, np.ndarray) and ids.dtype == np.int32 and ids.flags.c_contiguous):
            raise B200RTError(E_INVALID, "borrow_ids needs a C-contiguous int32 array (a converted copy would be freed too early)")
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if ids.ndim != 2:
            raise B200RTError(E_INVALID, f"ids must be [n_items, max_len], got shape {ids.shape}")
        n, S = ids.shape
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.int32)
            if lens.shape != (n,):
                raise B200RTError(E_INVALID, f"lens must have shape ({n},), got {lens.shape}")
        if out is None:
            out = np.empty((n, self.hidden), np.float32)
        elif out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != (n, self.hidden):
            raise B200RTError(E_INVALID, "out must be a C-contiguous float32 [n_items, hidden] array")
        t = ctypes.c_uint64(0)
        with _live_lock:
            _check(self._lib.b200rt_submit_ex(self.handle, _ptr(ids), _ptr(lens) if lens is not None else None, n, S, _ptr(out),
                                              SUBMIT_BORROW_IDS if borrow_ids else 0, ctypes.byref(t)))
            tk = Ticket(t.value, out, tag, ids if borrow_ids else None)
            _live[tk.id] = tk
        return tk

    def wait(self, ticket: Ticket, timeout_ms: int = -1) -> np.ndarray | None:
        rc = self._lib.b200rt_wait(ticket.id, timeout_ms)
        if rc == TIMEOUT:
            return None
        with _live_lock:
            _live.pop(ticket.id, None)
        if rc < 0:
            raise TicketError(rc, (self._lib.b200rt_last_error() or b"").decode(errors="replace"), ticket)
        return ticket.out

    def poll_any(self, timeout_ms: int = -1) -> Ticket | None:
        """Process-wide (see :func:`poll_any`): may return a ticket submitted through another model."""
        return poll_any(timeout_ms)

    def embed(self, ids: np.ndarray, lens=None) -> np.ndarray:
        """Synchronous convenience: one input in, its embeddings out.  The caller is blocked for the duration, so a
        C-contiguous int32 ``ids`` is lent to the library instead of copied."""
        lend = isinstance(ids, np.ndarray) and ids.dtype == np.int32 and ids.flags.c_contiguous
        return self.wait(self.submit(ids, lens, borrow_ids=lend))

    def embed_device(self, gpu: int, d_ids: int, d_lens: int, n_items: int, max_len: int, d_out: int, stream: int = 0):
        """Device-resident forward (raw device addresses); asynchronous on ``stream``."""
        _check(self._lib.b200rt_embed_device(self.handle, gpu, d_ids, d_lens, n_items, max_len, d_out, stream or None))

    # ---- debug helpers (tests only)
    def debug_hidden(self, ids: np.ndarray, lens, n_layers: int) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n, S = ids.shape
        lens_a = np.ascontiguousarray(lens, dtype=np.int32) if lens is not None else None
        out = np.empty((n * S, self.hidden), np.float32)
        _check(self._lib.b200rt_debug_hidden(self.handle, _ptr(ids), _ptr(lens_a) if lens_a is not None else None, n, S, n_layers, _ptr(out)))
        return out.reshape(n, S, self.hidden)

    def profile_forward(self, n_items: int, max_len: int, iters: int = 3) -> dict:
        names = ctypes.create_string_buffer(4096)
        ms = (ctypes.c_float * 64)()
        n = ctypes.c_int32(0)
        _check(self._lib.b200rt_debug_profile_forward(self.handle, n_items, max_len, iters, names, 4096, ms, ctypes.byref(n), 64))
        parts = names.raw.split(b"\0")
        return {parts[i].decode(): float(ms[i]) for i in range(n.value)}


class ImageEmbedModel:
    """A loaded CLIP-style ViT image tower (kind "vit"): ``submit``/``wait`` take preprocessed pixels
    ``[n, 3, image, image]`` float32 and give ``[n, proj]`` unit-norm embeddings -- the in-box stand-in for
    ``engine.image_embed(images=...)`` in the reference's image example (image_embeddings_infinity.py:330-350)."""

    def __init__(self, geometry: dict, blob: np.ndarray):
        lib = load_library()
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        g = geometry
        self.cfg = VitConfig(g["image"], g["patch"], g["hidden"], g["layers"], g["heads"], g["inter"], g["proj"], g["eps"])
        self.image, self.proj = g["image"], g["proj"]
        self.tokens = (g["image"] // g["patch"]) ** 2 + 1
        self.hidden = g["hidden"]
        h = ctypes.c_int32(-1)
        _check(lib.b200rt_model_load(b"vit", ctypes.byref(self.cfg), _ptr(blob), blob.nbytes, ctypes.byref(h)))
        self.handle = h.value
        self._lib = lib

    def submit(self, pixels: np.ndarray, out: np.ndarray | None = None, tag=None) -> Ticket:
        """``pixels`` (C-contiguous float32) and ``out`` stay lent to the library until the ticket completes."""
        if not (isinstance(pixels, np.ndarray) and pixels.dtype == np.float32 and pixels.flags.c_contiguous):
            pixels = np.ascontiguousarray(pixels, dtype=np.float32)
        if pixels.ndim != 4 or pixels.shape[1:] != (3, self.image, self.image):
            raise B200RTError(E_INVALID, f"pixels must be [n, 3, {self.image}, {self.image}], got shape {pixels.shape}")
        n = pixels.shape[0]
        if out is None:
            out = np.empty((n, self.proj), np.float32)
        elif out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != (n, self.proj):
            raise B200RTError(E_INVALID, "out must be a C-contiguous float32 [n_items, proj] array")
        t = ctypes.c_uint64(0)
        with _live_lock:
            _check(self._lib.b200rt_submit_pixels(self.handle, _ptr(pixels), n, _ptr(out), ctypes.byref(t)))
            tk = Ticket(t.value, out, tag, pixels)
            _live[tk.id] = tk
        return tk

    wait = EmbedModel.wait
    poll_any = EmbedModel.poll_any

    def embed(self, pixels: np.ndarray) -> np.ndarray:
        return self.wait(self.submit(pixels))

    def debug_hidden(self, pixels: np.ndarray, n_layers: int) -> np.ndarray:
        pixels = np.ascontiguousarray(pixels, dtype=np.float32)
        n = pixels.shape[0]
        out = np.empty((n * self.tokens, self.hidden), np.float32)
        _check(self._lib.b200rt_debug_vit_hidden(self.handle, _ptr(pixels), n, n_layers, _ptr(out)))
        return out.reshape(n, self.tokens, self.hidden)


def device_sync(gpu: int = 0):
    _check(load_library().b200rt_device_sync(gpu))


def debug_gemm(epi: int, a16: np.ndarray, w16: np.ndarray, bias: np.ndarray, resid=None, iters: int = 1, ln_stats=None,
               ln_gamma=None, ln_beta=None, eps: float = 1e-12, want_stats: bool = False):
    """a16 [M,K], w16 [N,K] float16; returns (out, ms) -- or (out, ms, stats_out [M, N/128, 2]) with ``want_stats`` (epi 2).
    ``ln_stats`` [M, parts, 2] (sum, M2) partials switch the LayerNorm epilogues on (see include/b200rt_debug.h)."""
    a16 = np.ascontiguousarray(a16, dtype=np.float16)
    w16 = np.ascontiguousarray(w16, dtype=np.float16)
    bias = np.ascontiguousarray(bias, dtype=np.float32)
    M, K = a16.shape
    N = w16.shape[0]
    out = np.empty((M, N), np.float32 if epi == 2 else np.float16)
    c32 = lambda v: np.ascontiguousarray(v, dtype=np.float32) if v is not None else None  # noqa: E731
    r, st, lg, lb = c32(resid), c32(ln_stats), c32(ln_gamma), c32(ln_beta)
    so = np.zeros((M, N // 128, 2), np.float32) if want_stats else None
    p = lambda v: _ptr(v) if v is not None else None  # noqa: E731
    ms = ctypes.c_float(0)
    _check(load_library().b200rt_debug_gemm(epi, _ptr(a16), _ptr(w16), _ptr(bias), p(r), _ptr(out), M, N, K, iters, ctypes.byref(ms),
                                            p(st), p(lg), p(lb), eps, p(so)))
    return (out, ms.value, so) if want_stats else (out, ms.value)


def debug_attention(qkv16: np.ndarray, lens: np.ndarray, B: int, S: int, iters: int = 1):
    qkv16 = np.ascontiguousarray(qkv16, dtype=np.float16)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    ctx = np.empty((B * S, 768), np.float16)
    ms = ctypes.c_float(0)
    _check(load_library().b200rt_debug_attention(_ptr(qkv16), _ptr(lens), _ptr(ctx), B, S, iters, ctypes.byref(ms)))
    return ctx, ms.value
