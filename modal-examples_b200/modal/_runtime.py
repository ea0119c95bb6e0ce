"""In-box execution core of the ``modal`` shim: where the reference SDK pickles a call, ships it over gRPC
and schedules a cloud container, this runs the call on a local worker -- a thread for sync callables, a
per-"container" asyncio loop for ``async def`` ones -- and leaves GPU work to ``libb200rt``'s C++ scheduler,
which the callables reach through ``b200rt``.  ``modal.is_local()`` is True in the driver and False inside
a worker, as the reference scripts expect.
"""
from __future__ import annotations

import asyncio
import collections
import queue
import concurrent.futures as cf
import contextvars
import inspect
import itertools
import os
import threading
import time
import uuid

# The map pump is many threads that each hold the GIL for microseconds between long GIL-free waits inside libb200rt.  With
# CPython's default 5 ms switch interval a worker returning from the C ABI can sit behind another thread's Python work for
# milliseconds -- longer than a wave takes on 8 GPUs.  The in-box runtime owns the process, so it shortens the interval.
import sys as _sys

if _sys.getswitchinterval() > 5e-4:
    _sys.setswitchinterval(5e-4)

_in_worker = contextvars.ContextVar("modal_shim_in_worker", default=False)
_task_ids = itertools.count(1)


def is_local() -> bool:
    return not _in_worker.get()


class _Loop:
    """A daemon thread running one asyncio loop: the in-box stand-in for one container's event loop."""

    def __init__(self, name):
        self.loop = asyncio.new_event_loop()
        self.thread = threading.Thread(target=self._run, name=name, daemon=True)
        self.thread.start()

    def _run(self):
        asyncio.set_event_loop(self.loop)
        self.loop.run_forever()

    def submit(self, coro) -> cf.Future:
        return asyncio.run_coroutine_threadsafe(coro, self.loop)

    def stop(self):
        self.loop.call_soon_threadsafe(self.loop.stop)


class Executor:
    """Bounded worker pool for one Function / one Cls instance."""

    def __init__(self, name: str, concurrency: int, env: dict | None = None):
        self.name = name
        self.concurrency = max(1, min(int(concurrency), 256))
        self.env = dict(env or {})
        self._pool = None
        self._loop = None
        self._sem = None
        self._lock = threading.Lock()

    @property
    def pool(self) -> cf.ThreadPoolExecutor:
        with self._lock:
            if self._pool is None:
                self._pool = cf.ThreadPoolExecutor(self.concurrency, thread_name_prefix=f"modal-{self.name}")
            return self._pool

    @property
    def loop(self) -> _Loop:
        with self._lock:
            if self._loop is None:
                self._loop = _Loop(f"modal-loop-{self.name}")
            return self._loop

    def _apply_env(self):
        for k, v in self.env.items():
            os.environ.setdefault(str(k), str(v))
        os.environ.setdefault("MODAL_TASK_ID", f"ta-local-{next(_task_ids):06d}")

    def submit(self, fn, args, kwargs, retries: int = 0) -> cf.Future:
        """Run fn(*args, **kwargs) on a worker; the returned Future carries the value or the callee's own
        exception object (original type and message, as FunctionCall.gather needs)."""
        self._apply_env()
        if inspect.iscoroutinefunction(fn):
            async def run_async():
                tok = _in_worker.set(True)
                try:
                    if self._sem is None:
                        self._sem = asyncio.Semaphore(self.concurrency)
                    async with self._sem:
                        for attempt in range(retries + 1):
                            try:
                                return await fn(*args, **kwargs)
                            except Exception:
                                if attempt == retries:
                                    raise
                finally:
                    _in_worker.reset(tok)

            return self.loop.submit(run_async())

        def run_sync():
            tok = _in_worker.set(True)
            try:
                for attempt in range(retries + 1):
                    try:
                        return fn(*args, **kwargs)
                    except Exception:
                        if attempt == retries:
                            raise
            finally:
                _in_worker.reset(tok)

        return self.pool.submit(run_sync)

    def shutdown(self):
        if self._pool is not None:
            self._pool.shutdown(wait=False, cancel_futures=True)
        if self._loop is not None:
            self._loop.stop()


def run_maybe_async(fn, *args, **kwargs):
    """Call a lifecycle hook that may be sync or ``async def`` (reference has both:
    06_gpu_and_ml/embeddings/image_embeddings_infinity.py:288-310)."""
    out = fn(*args, **kwargs)
    if inspect.isawaitable(out):
        try:
            asyncio.get_running_loop()
        except RuntimeError:
            return asyncio.run(out)
        # already inside a loop (async entrypoint): run on a helper thread
        with cf.ThreadPoolExecutor(1) as ex:
            return ex.submit(asyncio.run, out).result()
    return out


# ------------------------------------------------------------------------------------------ map engine


def map_sync(submit_one, inputs, window: int, order_outputs: bool, return_exceptions: bool):
    """Generator over results.  ``inputs`` is consumed lazily from the caller's thread with at most
    ``window`` calls in flight (back-pressure on generator inputs such as
    text_embeddings_inference.py:156-167's ``generate_batches()``).

    Unordered delivery costs O(1) per result: every future pushes itself onto a completion queue when it finishes.  (Waiting
    on the whole window with ``concurrent.futures.wait`` is O(window) lock traffic per result -- with ~200 calls in flight
    that alone capped the pump near 2 k inputs/s, the rate an 8-GPU pool needs.)"""
    it = iter(inputs)
    exhausted = False
    pending = collections.deque() if order_outputs else set()  # futures in flight (submission order when ordered)
    done_q = queue.SimpleQueue()

    def refill():
        nonlocal exhausted
        while not exhausted and len(pending) < window:
            try:
                a = next(it)
            except StopIteration:
                exhausted = True
                break
            fut = submit_one(a)
            if order_outputs:
                pending.append(fut)
            else:
                pending.add(fut)
                fut.add_done_callback(done_q.put)

    def deliver(fut):
        exc = fut.exception()
        if exc is None:
            return fut.result()
        if return_exceptions:
            return exc
        for f in list(pending):
            f.cancel()
        raise exc

    refill()
    while pending:
        if order_outputs:
            fut = pending.popleft()
            fut.exception()  # wait
        else:
            fut = done_q.get()
            pending.discard(fut)
        val = deliver(fut)
        refill()
        yield val


async def map_async(submit_one, inputs, window: int, order_outputs: bool, return_exceptions: bool):
    """``.map.aio``: async generator twin of :func:`map_sync`; accepts sync or async iterables."""
    loop = asyncio.get_running_loop()
    if hasattr(inputs, "__aiter__"):
        ait = inputs.__aiter__()

        async def nxt():
            try:
                return True, await ait.__anext__()
            except StopAsyncIteration:
                return False, None
    else:
        sit = iter(inputs)

        async def nxt():
            try:
                return True, next(sit)
            except StopIteration:
                return False, None

    pending = collections.deque() if order_outputs else set()
    done_q: asyncio.Queue = asyncio.Queue()
    exhausted = False

    async def refill():
        nonlocal exhausted
        while not exhausted and len(pending) < window:
            ok, a = await nxt()
            if not ok:
                exhausted = True
                break
            fut = asyncio.wrap_future(submit_one(a), loop=loop)
            if order_outputs:
                pending.append(fut)
            else:
                pending.add(fut)
                fut.add_done_callback(done_q.put_nowait)  # runs on this loop: O(1) per result instead of asyncio.wait over the window

    await refill()
    while pending:
        if order_outputs:
            fut = pending.popleft()
            await asyncio.wait([fut])
        else:
            fut = await done_q.get()
            pending.discard(fut)
        exc = None if fut.cancelled() else fut.exception()
        if fut.cancelled():
            exc = asyncio.CancelledError()
        if exc is not None and not return_exceptions:
            for f in list(pending):
                f.cancel()
            raise exc
        await refill()
        yield exc if exc is not None else fut.result()


# ------------------------------------------------------------------------------------------ dynamic batching


class Batcher:
    """``@modal.batched(max_batch_size, wait_ms)``: callers pass single items, the function is invoked once
    with lists and must return an equal-length list (reference 03_scaling_out/dynamic_batching.py:28-60,
    06_gpu_and_ml/speech-to-text/batched_whisper.py:127-138)."""

    def __init__(self, fn, max_batch_size: int, wait_ms: int, bound_self=None):
        self.fn, self.max_batch_size, self.wait_s, self.bound_self = fn, int(max_batch_size), wait_ms / 1000.0, bound_self
        self.q = collections.deque()
        self.cv = threading.Condition()
        self.thread = threading.Thread(target=self._run, daemon=True, name=f"modal-batcher-{getattr(fn, '__name__', 'fn')}")
        self.thread.start()

    def submit(self, args, kwargs) -> cf.Future:
        fut = cf.Future()
        with self.cv:
            self.q.append((args, kwargs, fut))
            self.cv.notify()
        return fut

    def _run(self):
        while True:
            with self.cv:
                while not self.q:
                    self.cv.wait()
                deadline = time.monotonic() + self.wait_s
                while len(self.q) < self.max_batch_size:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self.cv.wait(left)
                batch = [self.q.popleft() for _ in range(min(len(self.q), self.max_batch_size))]
            try:
                sig = inspect.signature(self.fn)
                names = [p for p in sig.parameters if p != "self"]
                cols = {}
                for args, kwargs, _ in batch:
                    bound = dict(zip(names, args))
                    bound.update(kwargs)
                    for k in names:
                        if k in bound:
                            cols.setdefault(k, []).append(bound[k])
                call_args = (self.bound_self,) if self.bound_self is not None else ()
                tok = _in_worker.set(True)
                try:
                    out = run_maybe_async(self.fn, *call_args, **cols)
                finally:
                    _in_worker.reset(tok)
                if len(out) != len(batch):
                    raise ValueError(f"batched function returned {len(out)} results for {len(batch)} inputs")
                for (_, _, fut), o in zip(batch, out):
                    fut.set_result(o)
            except Exception as e:  # noqa: BLE001
                for _, _, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)


def new_object_id(prefix: str) -> str:
    return f"{prefix}-{uuid.uuid4().hex[:22]}"
