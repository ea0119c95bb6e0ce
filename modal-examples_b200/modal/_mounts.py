"""Virtual volume mounts: ``volumes={"/data": vol}`` (reference: text_embeddings_inference.py:27, 141-145) without
touching ``/``.  The reference scripts address their volumes through hard-coded absolute paths (``DATA_PATH =
Path("/data/dataset.jsonl")``); in a container those are real mount points, in-box they would be paths outside the
runtime's state directory.  Instead of creating them, path arguments of the handful of filesystem entry points the
scripts use (``open``, ``os.stat`` -- hence ``Path.exists`` --, ``os.listdir``/``scandir``, ``os.makedirs`` ...) are
rewritten when they fall under a registered mount point that does not exist on the real filesystem.

``MODAL_SHIM_MOUNTS=virtual`` (default) | ``link`` (symlink the mount point, needs write access to its parent; the
round-1 behaviour behind ``MODAL_SHIM_LINK_MOUNTS=1``) | ``off``.
"""
from __future__ import annotations

import builtins
import io
import os
import threading

_lock = threading.Lock()
_mounts: dict[str, str] = {}  # normalised mount point -> local directory
_installed = False
_orig = {}


def mode() -> str:
    if os.environ.get("MODAL_SHIM_LINK_MOUNTS") == "1":
        return "link"
    return os.environ.get("MODAL_SHIM_MOUNTS", "virtual")


def translate(path):
    """The real location of ``path`` if it lies under a virtual mount, else ``path`` unchanged (same type family)."""
    if not _mounts or isinstance(path, int):
        return path
    try:
        p = os.fspath(path)
    except TypeError:
        return path
    if isinstance(p, bytes):
        return path
    if not p.startswith("/"):
        return path
    norm = os.path.normpath(p)
    for mp, local in _mounts.items():
        if norm == mp or norm.startswith(mp + "/"):
            return local + norm[len(mp):]
    return path


def _wrap(fn):
    def wrapper(path, *a, **k):
        return fn(translate(path), *a, **k)

    wrapper.__name__ = getattr(fn, "__name__", "wrapped")
    wrapper.__doc__ = getattr(fn, "__doc__", None)
    wrapper.__wrapped__ = fn
    return wrapper


def _wrap2(fn):
    def wrapper(src, dst, *a, **k):
        return fn(translate(src), translate(dst), *a, **k)

    wrapper.__wrapped__ = fn
    return wrapper


def _install():
    global _installed
    if _installed:
        return
    _orig["open"] = builtins.open
    opened = _wrap(builtins.open)
    builtins.open = opened
    io.open = opened
    for name in ("stat", "lstat", "listdir", "scandir", "mkdir", "makedirs", "remove", "unlink", "rmdir", "access", "chdir", "utime"):
        if hasattr(os, name):
            _orig["os." + name] = getattr(os, name)
            setattr(os, name, _wrap(getattr(os, name)))
    for name in ("rename", "replace"):
        _orig["os." + name] = getattr(os, name)
        setattr(os, name, _wrap2(getattr(os, name)))
    _installed = True


# Path-taking entry points of C extensions that bypass ``open``/``os`` and that the reference's scripts call with volume
# paths: torchvision's file reader / writer behind ``read_image`` / ``write_jpeg`` (image_embeddings_infinity.py:176-186,
# 318-320).  Wrapped when their module is loaded by the time a mount is registered (scripts import them at module level).
_FOREIGN = (("torchvision.io.image", "read_file", 0), ("torchvision.io.image", "write_file", 0))


def _patch_foreign():
    import sys

    for mod_name, attr, _pos in _FOREIGN:
        mod = sys.modules.get(mod_name)
        fn = getattr(mod, attr, None) if mod is not None else None
        if fn is not None and not hasattr(fn, "__wrapped__"):
            _orig[f"{mod_name}.{attr}"] = fn
            setattr(mod, attr, _wrap(fn))


def register(mount_point, local_dir: str) -> bool:
    """Make ``mount_point`` resolve to ``local_dir``.  Returns whether the mount point is usable afterwards."""
    mp = os.path.normpath(str(mount_point))
    m = mode()
    if m == "off":
        return os.path.exists(mp)
    real_exists = (_orig.get("os.stat") or os.stat)
    try:
        real_exists(mp)
        exists = True
    except OSError:
        exists = False
    if exists or os.path.islink(mp):
        return True  # a real directory (or an earlier link) is there: leave the filesystem in charge
    if m == "link":
        try:
            os.symlink(local_dir, mp)
            return True
        except OSError:
            return False
    if not mp.startswith("/") or mp == "/":
        return False
    with _lock:
        _mounts[mp] = local_dir
        _install()
        _patch_foreign()
    return True


def registered() -> dict:
    return dict(_mounts)
