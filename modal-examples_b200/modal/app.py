"""``modal.App``: registry + decorators.  ``@app.function`` / ``@app.cls`` / ``@app.local_entrypoint`` with the
keyword census of SURVEY.md §2.2; unknown keywords are accepted and recorded (cloud-only knobs have no in-box
meaning), ``gpu=`` is parsed to a local B200 count."""
from __future__ import annotations

import contextlib
import inspect

from . import _runtime as rt
from .cls import Cls, marks
from .functions import Function
from .gpu import parse_gpu_count

_apps: dict[str, "App"] = {}


def _lookup_app(name):
    return _apps.get(name)


class LocalEntrypoint:
    def __init__(self, raw_f, app):
        self.raw_f, self.app = raw_f, app
        self.__name__ = raw_f.__name__
        self.__doc__ = raw_f.__doc__

    def __call__(self, *a, **k):
        with self.app.run():
            out = self.raw_f(*a, **k)
            if inspect.isawaitable(out):
                return rt.run_maybe_async(lambda: out)
            return out


class App:
    def __init__(self, name=None, *, image=None, secrets=None, volumes=None, **kwargs):
        self.name = name
        self.description = name
        self.image, self.secrets, self.volumes = image, list(secrets or []), dict(volumes or {})
        self.registered_functions: dict[str, Function] = {}
        self.registered_classes: dict[str, Cls] = {}
        self.registered_entrypoints: dict[str, LocalEntrypoint] = {}
        self.registered_web_endpoints: list[str] = []
        self.app_id = rt.new_object_id("ap")
        self._running = 0
        if name:
            _apps[name] = self

    # ---- decorators
    def _options(self, kw):
        o = dict(kw)
        o.setdefault("image", self.image)
        o["gpu_count"] = parse_gpu_count(o.get("gpu"))
        vols = dict(self.volumes)
        vols.update(o.get("volumes") or {})
        o["volumes"] = vols
        for path, vol in vols.items():
            if hasattr(vol, "mount_at"):
                vol.mount_at(path)
        return o

    def function(self, _fn=None, **kwargs):
        def deco(fn):
            if isinstance(fn, Function):
                fn.options.update(self._options(kwargs))
                return fn
            opts = self._options(kwargs)
            m = marks(fn)
            if "max_inputs" in m:
                opts["max_inputs"] = m["max_inputs"]
            batcher = None
            if "batched" in m:
                holder = {}

                def batcher():
                    if "b" not in holder:
                        holder["b"] = rt.Batcher(fn, m["batched"][0], m["batched"][1])
                    return holder["b"]
            f = Function(fn, tag=kwargs.get("name") or fn.__name__, app=self, options=opts, batcher=batcher)
            self.registered_functions[f.tag] = f
            if "web" in m:
                self.registered_web_endpoints.append(f.tag)
            return f

        return deco(_fn) if callable(_fn) else deco

    def cls(self, _cls=None, **kwargs):
        def deco(user_cls):
            c = Cls(user_cls, self, self._options(kwargs))
            self.registered_classes[user_cls.__name__] = c
            return c

        return deco(_cls) if inspect.isclass(_cls) else deco

    def local_entrypoint(self, _fn=None, **kwargs):
        def deco(fn):
            e = LocalEntrypoint(fn, self)
            self.registered_entrypoints[fn.__name__] = e
            return e

        return deco(_fn) if callable(_fn) else deco

    # ---- lifecycle
    @contextlib.contextmanager
    def run(self, **_kw):
        self._running += 1
        try:
            yield self
        finally:
            self._running -= 1
            if self._running == 0:
                from .cls import _shutdown_all

                _shutdown_all()

    def deploy(self, **_kw):
        if self.name:
            _apps[self.name] = self
        return self

    def include(self, other: "App"):
        self.registered_functions.update(other.registered_functions)
        self.registered_classes.update(other.registered_classes)
        return self

    @staticmethod
    def lookup(name, *, create_if_missing=False, **_kw) -> "App":
        if name in _apps:
            return _apps[name]
        if create_if_missing:
            return App(name)
        from .exception import NotFoundError

        raise NotFoundError(f"App {name!r} not found in this process")

    def __repr__(self):
        return f"<modal.App {self.name!r}>"


Stub = App  # legacy alias still present in the tree


@contextlib.contextmanager
def enable_output(*_a, **_k):
    yield
