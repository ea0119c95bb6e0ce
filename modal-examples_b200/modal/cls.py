"""``@app.cls`` lifecycle and the method / lifecycle decorators: a ``Cls()`` handle is lazy;
the first ``.method.remote/map/spawn`` builds the user object once per parameter set (the in-box "container"),
runs every ``@modal.enter`` (sync or async), and ``@modal.exit`` runs when the app context closes."""
from __future__ import annotations

import atexit
import threading

from . import _runtime as rt
from .exception import NotFoundError
from .functions import Function

_MARK = "_modal_shim"


def _mark(fn, **kv):
    d = dict(getattr(fn, _MARK, {}))
    d.update(kv)
    try:
        setattr(fn, _MARK, d)
    except AttributeError:
        pass
    return fn


def marks(fn) -> dict:
    return getattr(fn, _MARK, {})


def _decorator(**kv):
    def outer(*dargs, **dkw):
        if len(dargs) == 1 and callable(dargs[0]) and not dkw:  # bare @modal.method
            return _mark(dargs[0], **kv)
        return lambda fn: _mark(fn, **kv, **{k + "_kwargs": dkw for k in kv if dkw})

    return outer


method = _decorator(method=True)
exit = _decorator(exit=True)  # noqa: A001 - mirrors modal.exit


def enter(*dargs, snap: bool = False, **_kw):
    if len(dargs) == 1 and callable(dargs[0]):
        return _mark(dargs[0], enter=True, snap=False)
    return lambda fn: _mark(fn, enter=True, snap=snap)


def concurrent(*, max_inputs: int = 1, target_inputs: int | None = None):
    def deco(obj):
        if isinstance(obj, Function):
            obj.options["max_inputs"] = max_inputs
            obj._executor = None
            return obj
        return _mark(obj, max_inputs=max_inputs)

    return deco


def batched(*, max_batch_size: int, wait_ms: int):
    return lambda fn: _mark(fn, batched=(int(max_batch_size), int(wait_ms)))


class _Parameter:
    def __init__(self, default=None, init=True):
        self.default = default
        self.has_default = default is not None


_NO_DEFAULT = object()


def parameter(*, default=_NO_DEFAULT, init: bool = True):
    p = _Parameter()
    p.default = None if default is _NO_DEFAULT else default
    p.has_default = default is not _NO_DEFAULT
    return p


_live_objs: list["Obj"] = []


def _shutdown_all():
    for o in list(_live_objs):
        o._teardown()


atexit.register(_shutdown_all)


class Obj:
    """One parameterised instance handle: ``Model()`` or ``Model(size="small")``."""

    def __init__(self, cls: "Cls", args, kwargs):
        self._cls, self._args, self._kwargs = cls, args, kwargs
        self._inst = None
        self._lock = threading.Lock()
        self._executor = None
        self._batchers = {}

    def _executor_for(self) -> rt.Executor:
        if self._executor is None:
            o = self._cls.options
            user_marks = marks(self._cls.user_cls)
            max_inputs = user_marks.get("max_inputs") or o.get("max_inputs") or o.get("allow_concurrent_inputs") or 1
            # One in-box instance stands in for the whole container pool (the reference's containers each run @enter --
            # e.g. spawn a server on a fixed port -- so N instances in one box would collide).  A class that opted into
            # input concurrency (@modal.concurrent / max_inputs > 1) declared its methods safe to overlap on one `self`:
            # it gets max_containers x max_inputs calls in flight.  A class that did not is never entered concurrently
            # (the reference guarantees one input per container): its calls are serialised on the single instance.
            conc = (o.get("max_containers") or o.get("concurrency_limit") or 1) * max_inputs if max_inputs > 1 else 1
            env = dict(getattr(o.get("image"), "_env", {}) or {})
            for s in o.get("secrets") or []:
                env.update(getattr(s, "_env", {}))
            self._executor = rt.Executor(self._cls.user_cls.__name__, conc, env)
        return self._executor

    def _instance(self):
        with self._lock:
            if self._inst is None:
                ucls = self._cls.user_cls
                params = {k: v for k, v in vars(ucls).items() if isinstance(v, _Parameter)}
                for klass in ucls.__mro__[1:]:
                    for k, v in vars(klass).items():
                        if isinstance(v, _Parameter):
                            params.setdefault(k, v)
                if params:
                    inst = ucls.__new__(ucls)
                    unknown = set(self._kwargs) - set(params)
                    if unknown or self._args:
                        raise TypeError(f"{ucls.__name__}() got unexpected parameters {sorted(unknown)}")
                    for k, p in params.items():
                        if k in self._kwargs:
                            setattr(inst, k, self._kwargs[k])
                        elif p.has_default:
                            setattr(inst, k, p.default)
                        else:
                            raise TypeError(f"{ucls.__name__}() missing required parameter {k!r}")
                else:
                    inst = ucls(*self._args, **self._kwargs)
                self._executor_for()._apply_env()
                tok = rt._in_worker.set(True)
                try:
                    for name in self._cls._hooks("enter"):
                        rt.run_maybe_async(getattr(inst, name))
                finally:
                    rt._in_worker.reset(tok)
                self._inst = inst
                _live_objs.append(self)
            return self._inst

    def _teardown(self):
        with self._lock:
            inst, self._inst = self._inst, None
        if self in _live_objs:
            _live_objs.remove(self)
        if inst is not None:
            for name in self._cls._hooks("exit"):
                try:
                    rt.run_maybe_async(getattr(inst, name))
                except TypeError:
                    rt.run_maybe_async(getattr(inst, name), None, None, None)
                except Exception as e:  # noqa: BLE001
                    print(f"[modal shim] @exit {name} raised {e!r}")
        if self._executor is not None:
            self._executor.shutdown()

    def __getattr__(self, name):
        ucls = self._cls.user_cls
        attr = getattr(ucls, name, None)
        if attr is None or not callable(attr):
            if attr is not None:
                return getattr(self._instance(), name)
            raise AttributeError(f"{ucls.__name__!r} has no attribute {name!r}")
        m = marks(attr)
        batcher = None
        if "batched" in m:
            def batcher(name=name, attr=attr, m=m):
                if name not in self._batchers:
                    self._batchers[name] = rt.Batcher(attr, m["batched"][0], m["batched"][1], bound_self=self._instance())
                return self._batchers[name]
        fn = Function(attr, tag=f"{ucls.__name__}.{name}", app=self._cls.app, executor=self._executor_for(), options=self._cls.options,
                      bound_self_factory=self._instance, batcher=batcher)
        self.__dict__[name] = fn
        return fn


class Cls:
    """Result of ``@app.cls(...)``."""

    def __init__(self, user_cls, app=None, options=None):
        self.user_cls, self.app, self.options = user_cls, app, dict(options or {})
        self.__name__ = user_cls.__name__
        self.__doc__ = user_cls.__doc__
        self._default = None

    def _hooks(self, kind):
        names = []
        for klass in reversed(self.user_cls.__mro__):
            for k, v in vars(klass).items():
                if callable(v) and marks(v).get(kind) and k not in names:
                    names.append(k)
        return names

    def __call__(self, *args, **kwargs) -> Obj:
        if not args and not kwargs:
            if self._default is None:
                self._default = Obj(self, (), {})
            return self._default
        return Obj(self, args, kwargs)

    def with_options(self, **opts) -> "Cls":
        o = dict(self.options)
        o.update(opts)
        return Cls(self.user_cls, self.app, o)

    def with_concurrency(self, *, max_inputs, target_inputs=None) -> "Cls":
        return self.with_options(max_inputs=max_inputs)

    def with_batching(self, **_k) -> "Cls":
        return self

    @staticmethod
    def from_name(app_name: str, name: str, **_kw) -> "Cls":
        from .app import _lookup_app

        app = _lookup_app(app_name)
        if app is None or name not in app.registered_classes:
            raise NotFoundError(f"Cls {app_name!r}/{name!r} is not deployed in this process (in-box runtime: import the app's module first)")
        return app.registered_classes[name]

    lookup = from_name

    def __getattr__(self, name):
        # class-level access to plain attributes / constants of the user class
        return getattr(self.user_cls, name)
