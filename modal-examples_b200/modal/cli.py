"""``modal run|serve|deploy`` for the in-box runtime.  Accepts the frontmatter ``cmd:`` forms the reference uses
(internal/utils.py:117-123): ``modal run file.py``, ``modal run file.py::fn``, ``modal run -m pkg.mod``,
``--detach``; entrypoint parameters become ``--kebab-case`` options (amazon_embeddings.py:3,50-55)."""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import inspect
import os
import sys
import typing

from . import _runtime as rt
from .app import App, LocalEntrypoint
from .functions import Function


def _import_target(ref: str, as_module: bool):
    path, _, fn_name = ref.partition("::")
    if as_module:
        mod = importlib.import_module(path)
    else:
        path = os.path.abspath(path)
        if not os.path.exists(path):
            raise SystemExit(f"modal: no such file {path}")
        sys.path.insert(0, os.path.dirname(path))
        spec = importlib.util.spec_from_file_location(os.path.splitext(os.path.basename(path))[0], path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
    return mod, fn_name or None


def _find_app(mod) -> App:
    for name in ("app", "stub"):
        if isinstance(getattr(mod, name, None), App):
            return getattr(mod, name)
    apps = [v for v in vars(mod).values() if isinstance(v, App)]
    if not apps:
        raise SystemExit("modal: no modal.App found in the module")
    return apps[0]


def _pick(app: App, mod, fn_name):
    if fn_name:
        obj = getattr(mod, fn_name, None) or app.registered_entrypoints.get(fn_name) or app.registered_functions.get(fn_name)
        if obj is None:
            raise SystemExit(f"modal: {fn_name!r} not found")
        return obj
    if len(app.registered_entrypoints) == 1:
        return next(iter(app.registered_entrypoints.values()))
    if not app.registered_entrypoints and len(app.registered_functions) == 1:
        return next(iter(app.registered_functions.values()))
    raise SystemExit("modal: specify which entrypoint to run with file.py::name "
                     f"(entrypoints: {sorted(app.registered_entrypoints)}, functions: {sorted(app.registered_functions)})")


def _strip_optional(t):
    if typing.get_origin(t) in (typing.Union, getattr(__import__("types"), "UnionType", None)):
        args = [a for a in typing.get_args(t) if a is not type(None)]
        if len(args) == 1:
            return args[0]
    return t


def _parse_fn_args(raw_f, argv):
    sig = inspect.signature(raw_f)
    try:
        hints = typing.get_type_hints(raw_f)
    except Exception:
        hints = {}
    ap = argparse.ArgumentParser(prog=f"modal run ...::{raw_f.__name__}")
    for name, p in sig.parameters.items():
        flag = "--" + name.replace("_", "-")
        t = _strip_optional(hints.get(name, p.annotation))
        if t is inspect.Parameter.empty:
            t = type(p.default) if p.default not in (inspect.Parameter.empty, None) else str
        required = p.default is inspect.Parameter.empty
        if t is bool:
            ap.add_argument(flag, action=argparse.BooleanOptionalAction, default=None if required else p.default, required=required)
        else:
            conv = t if t in (int, float, str) else str
            ap.add_argument(flag, type=conv, default=None if required else p.default, required=required)
    ns = ap.parse_args(argv)
    return {k: getattr(ns, k) for k in sig.parameters}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print("usage: modal {run|serve|deploy} [-m] [--detach] <file.py|module>[::function] [--fn-arg value ...]")
        return 0
    verb, rest = argv[0], argv[1:]
    if verb not in ("run", "serve", "deploy", "shell"):
        raise SystemExit(f"modal: unsupported command {verb!r} in the in-box runtime")
    as_module = False
    flags = []
    while rest and rest[0].startswith("-"):
        f = rest.pop(0)
        if f == "-m":
            as_module = True
        elif f in ("-e", "--env", "--name", "--tag", "--timeout"):
            flags.append((f, rest.pop(0)))
        else:
            flags.append((f, None))  # --detach, -q, -i, --stream-logs ...
    if not rest:
        raise SystemExit("modal: missing target")
    mod, fn_name = _import_target(rest[0], as_module)
    app = _find_app(mod)
    if verb == "deploy":
        app.deploy(_source=getattr(mod, "__file__", None))
        print(f"[modal b200] deployed app {app.name!r} ({len(app.registered_functions)} functions, {len(app.registered_classes)} classes)")
        return 0
    if verb == "serve":
        print(f"[modal b200] serve: web endpoints are out of scope in-box; app {app.name!r} imported OK "
              f"({len(app.registered_web_endpoints)} endpoints recorded)")
        return 0
    target = _pick(app, mod, fn_name)
    raw = target.raw_f if isinstance(target, (LocalEntrypoint, Function)) else target
    kwargs = _parse_fn_args(raw, rest[1:])
    with app.run():
        if isinstance(target, Function):
            out = target.remote(**kwargs)  # `modal run file.py::fn` on an @app.function runs it as a remote call
        else:
            out = raw(**kwargs)
            if inspect.isawaitable(out):
                out = rt.run_maybe_async(lambda: out)
        if any(f == "--detach" or f == "-d" for f, _ in flags):
            from .functions import drain_spawned

            n = drain_spawned()
            if n:
                print(f"[modal b200] --detach: waited for {n} spawned call(s) to finish")
    if out is not None and isinstance(target, Function):
        print(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
