"""``modal.Function`` / ``modal.FunctionCall``: the invocation surface the reference scripts use
(census in SURVEY.md §2.2): ``.local .remote .map .starmap .for_each .spawn .remote_gen`` and the ``.aio``
twin of each."""
from __future__ import annotations

import asyncio
import concurrent.futures as cf
import inspect
import threading

from . import _runtime as rt
from .exception import NotFoundError


class _Invokable:
    """A bound call style with a ``.aio`` twin: ``f.remote(x)`` and ``await f.remote.aio(x)``."""

    def __init__(self, sync_impl, aio_impl):
        self._sync, self.aio = sync_impl, aio_impl

    def __call__(self, *args, **kwargs):
        return self._sync(*args, **kwargs)


_calls: dict[str, "FunctionCall"] = {}
_calls_lock = threading.Lock()


class FunctionCall:
    """Handle returned by ``.spawn`` (reference 08_advanced/parallel_execution.py:33-48,
    poll_delayed_result.py:43-57, amazon_embeddings.py:104-116)."""

    def __init__(self, future: cf.Future):
        self._future = future
        self.object_id = rt.new_object_id("fc")
        with _calls_lock:
            _calls[self.object_id] = self

    def _get(self, timeout=None):
        try:
            return self._future.result(timeout)
        except cf.TimeoutError:
            raise TimeoutError(f"FunctionCall {self.object_id} not finished within {timeout}s") from None

    async def _get_aio(self, timeout=None):
        try:
            return await asyncio.wait_for(asyncio.wrap_future(self._future), timeout)
        except asyncio.TimeoutError:
            raise TimeoutError(f"FunctionCall {self.object_id} not finished within {timeout}s") from None

    @property
    def get(self):
        return _Invokable(self._get, self._get_aio)

    def cancel(self):
        self._future.cancel()

    @staticmethod
    def from_id(object_id: str) -> "FunctionCall":
        with _calls_lock:
            if object_id not in _calls:
                raise NotFoundError(f"FunctionCall {object_id!r} not found in this process")
            return _calls[object_id]

    @staticmethod
    def _gather(*calls):
        # first failure re-raises the callee's own exception (type + message preserved)
        return [c._get() for c in calls]

    @staticmethod
    async def _gather_aio(*calls):
        return list(await asyncio.gather(*(c._get_aio() for c in calls)))


FunctionCall.gather = _Invokable(FunctionCall._gather, FunctionCall._gather_aio)  # type: ignore[attr-defined]


def drain_spawned(timeout=None) -> int:
    """Block until every `.spawn`ed call of this process has finished (their outcome stays with the handle); returns how many
    were still running.  `modal run --detach` keeps an app alive until its spawned inputs are done
    (amazon_embeddings.py:2,17-18,104-116); in-box the CLI waits here before the containers are torn down."""
    with _calls_lock:
        pending = [c._future for c in _calls.values() if not c._future.done()]
    cf.wait(pending, timeout=timeout)
    return len(pending)


def _zip_inputs(iterables):
    if len(iterables) == 1:
        return ((x,) for x in iterables[0])
    return zip(*iterables)


async def _azip_inputs(iterables):
    if len(iterables) == 1 and hasattr(iterables[0], "__aiter__"):
        async for x in iterables[0]:
            yield (x,)
    else:
        for t in _zip_inputs(iterables):
            yield t


class Function:
    """An ``@app.function`` (or a bound ``@modal.method``).  ``raw_f`` runs in-box on an Executor worker."""

    def __init__(self, raw_f, *, tag=None, app=None, executor=None, options=None, bound_self_factory=None, batcher=None):
        self.raw_f = raw_f
        self.tag = tag or getattr(raw_f, "__name__", "function")
        self.app = app
        self.options = dict(options or {})
        self._executor = executor
        self._self_factory = bound_self_factory  # for methods: returns the live user object (runs @enter once)
        self._batcher = batcher
        self.object_id = rt.new_object_id("fu")
        self.is_generator = inspect.isgeneratorfunction(raw_f) or inspect.isasyncgenfunction(raw_f)
        self.__name__ = getattr(raw_f, "__name__", self.tag)
        self.__doc__ = getattr(raw_f, "__doc__", None)
        self.remote = _Invokable(self._remote, self._remote_aio)
        self.spawn = _Invokable(self._spawn, self._spawn_aio)
        self.map = _Invokable(self._map, self._map_aio)
        self.starmap = _Invokable(self._starmap, self._starmap_aio)
        self.for_each = _Invokable(self._for_each, self._for_each_aio)
        self.remote_gen = _Invokable(self._remote_gen, self._remote_gen_aio)

    # -- plumbing
    @property
    def executor(self) -> rt.Executor:
        if self._executor is None:
            o = self.options
            conc = (o.get("max_containers") or o.get("concurrency_limit") or 8) * (o.get("max_inputs") or o.get("allow_concurrent_inputs") or 1)
            env = {}
            img = o.get("image") or (self.app and getattr(self.app, "image", None))
            if img is not None:
                env.update(getattr(img, "_env", {}))
            for s in list(o.get("secrets") or []) + list(getattr(self.app, "secrets", None) or []):
                env.update(getattr(s, "_env", {}))
            self._executor = rt.Executor(self.tag, max(conc, 32 if not o.get("gpu") else conc), env)
        return self._executor

    def _retries(self) -> int:
        r = self.options.get("retries")
        return int(getattr(r, "max_retries", r) or 0)

    def _call_args(self, args):
        if self._self_factory is not None:
            return (self._self_factory(),) + tuple(args)
        return tuple(args)

    def _submit(self, args, kwargs) -> cf.Future:
        if self._batcher is not None:
            return self._batcher().submit(tuple(args), dict(kwargs))
        return self.executor.submit(self.raw_f, self._call_args(args), kwargs, self._retries())

    # -- call styles
    def local(self, *args, **kwargs):
        return self.raw_f(*self._call_args(args), **kwargs)

    def __call__(self, *args, **kwargs):
        # inside a worker a Function called directly behaves like the plain function (reference scripts do this)
        return self.local(*args, **kwargs)

    def _remote(self, *args, **kwargs):
        if self.is_generator:
            raise TypeError(f"{self.tag} is a generator function: use .remote_gen()")
        return self._submit(args, kwargs).result()

    async def _remote_aio(self, *args, **kwargs):
        return await asyncio.wrap_future(self._submit(args, kwargs))

    def _spawn(self, *args, **kwargs) -> FunctionCall:
        return FunctionCall(self._submit(args, kwargs))

    async def _spawn_aio(self, *args, **kwargs) -> FunctionCall:
        return FunctionCall(self._submit(args, kwargs))

    def _window(self):
        return max(2 * self.executor.concurrency, 8)

    def _map(self, *iterables, kwargs=None, order_outputs=True, return_exceptions=False, wrap_returned_exceptions=True,
             wrap_return_exceptions=None):
        kw = dict(kwargs or {})
        return rt.map_sync(lambda a: self._submit(a, kw), _zip_inputs(iterables), self._window(), order_outputs, return_exceptions)

    def _map_aio(self, *iterables, kwargs=None, order_outputs=True, return_exceptions=False, wrap_returned_exceptions=True,
                 wrap_return_exceptions=None):
        kw = dict(kwargs or {})
        return rt.map_async(lambda a: self._submit(a, kw), _azip_inputs(iterables), self._window(), order_outputs, return_exceptions)

    def _starmap(self, iterable, *, kwargs=None, order_outputs=True, return_exceptions=False, **_ignored):
        kw = dict(kwargs or {})
        return rt.map_sync(lambda a: self._submit(tuple(a), kw), iterable, self._window(), order_outputs, return_exceptions)

    def _starmap_aio(self, iterable, *, kwargs=None, order_outputs=True, return_exceptions=False, **_ignored):
        kw = dict(kwargs or {})
        return rt.map_async(lambda a: self._submit(tuple(a), kw), iterable, self._window(), order_outputs, return_exceptions)

    def _for_each(self, *iterables, kwargs=None, ignore_exceptions=False):
        for _ in self._map(*iterables, kwargs=kwargs, order_outputs=False, return_exceptions=ignore_exceptions):
            pass

    async def _for_each_aio(self, *iterables, kwargs=None, ignore_exceptions=False):
        async for _ in self._map_aio(*iterables, kwargs=kwargs, order_outputs=False, return_exceptions=ignore_exceptions):
            pass

    def _remote_gen(self, *args, **kwargs):
        # generator functions stream their values (reference 01_getting_started/generators.py:13-22)
        call_args = self._call_args(args)
        if inspect.isasyncgenfunction(self.raw_f):
            loop = self.executor.loop
            agen = self.raw_f(*call_args, **kwargs)
            while True:
                try:
                    yield loop.submit(agen.__anext__()).result()
                except StopAsyncIteration:
                    return
        # the body of a sync generator runs inside next(): is_local() must read False there too
        gen = None
        while True:
            tok = rt._in_worker.set(True)
            try:
                if gen is None:
                    gen = self.raw_f(*call_args, **kwargs)
                try:
                    v = next(gen)
                except StopIteration:
                    return
            finally:
                rt._in_worker.reset(tok)
            yield v

    async def _remote_gen_aio(self, *args, **kwargs):
        call_args = self._call_args(args)
        if inspect.isasyncgenfunction(self.raw_f):
            async for v in self.raw_f(*call_args, **kwargs):
                yield v
        else:
            for v in self.raw_f(*call_args, **kwargs):
                yield v
                await asyncio.sleep(0)

    # -- misc surface
    def get_web_url(self):
        return f"http://127.0.0.1:0/{self.tag}"

    web_url = property(lambda self: self.get_web_url())

    def keep_warm(self, *_a, **_k):
        return None

    class _Autoscaler:
        def __call__(self, **_k):
            return None

        async def aio(self, **_k):
            return None

    update_autoscaler = _Autoscaler()

    @staticmethod
    def from_name(app_name: str, name: str, **_kw) -> "Function":
        from .app import _lookup_app

        app = _lookup_app(app_name)
        if app is None or name not in app.registered_functions:
            raise NotFoundError(f"Function {app_name!r}/{name!r} is not registered in this process (in-box runtime: import the app's module first)")
        return app.registered_functions[name]

    lookup = from_name

    def __repr__(self):
        return f"<modal.Function {self.tag}>"
