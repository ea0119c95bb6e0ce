"""``modal.Image``: container builds have no in-box meaning, so every builder method is a chainable recorded
no-op.  Three behaviours matter: ``imports()`` swallows ImportError
(text_embeddings_inference.py:75-76), ``run_function`` must NOT execute (``:70`` would spawn TEI), and
``.env({...})`` is remembered and applied to the worker's environment."""
from __future__ import annotations

import contextlib


class Image:
    def __init__(self, steps=None, env=None):
        self._steps = list(steps or [])
        self._env = dict(env or {})

    def _with(self, name, *args, **kwargs) -> "Image":
        return Image(self._steps + [(name, args, kwargs)], self._env)

    # constructors
    @staticmethod
    def debian_slim(python_version=None, **kw):
        return Image([("debian_slim", (python_version,), kw)])

    @staticmethod
    def from_registry(tag, *a, **kw):
        return Image([("from_registry", (tag,) + a, kw)])

    @staticmethod
    def from_dockerfile(path, *a, **kw):
        return Image([("from_dockerfile", (path,) + a, kw)])

    @staticmethod
    def micromamba(python_version=None, **kw):
        return Image([("micromamba", (python_version,), kw)])

    @staticmethod
    def from_aws_ecr(tag, *a, **kw):
        return Image([("from_aws_ecr", (tag,) + a, kw)])

    @staticmethod
    def from_gcp_artifact_registry(tag, *a, **kw):
        return Image([("from_gcp_artifact_registry", (tag,) + a, kw)])

    def env(self, vars):  # noqa: A002
        e = dict(self._env)
        e.update({str(k): str(v) for k, v in dict(vars).items()})
        return Image(self._steps + [("env", (dict(vars),), {})], e)

    def run_function(self, fn, *a, **kw):
        return self._with("run_function", getattr(fn, "__name__", repr(fn)), *a, **kw)  # recorded, never executed

    @contextlib.contextmanager
    def imports(self):
        try:
            yield
        except ImportError:
            pass

    def __getattr__(self, name):
        # pip_install, uv_pip_install, apt_install, run_commands, add_local_dir, add_local_file, add_local_python_source,
        # dockerfile_commands, entrypoint, workdir, micromamba_install, poetry_install_from_file, pip_install_from_requirements, cmd ...
        if name.startswith("_"):
            raise AttributeError(name)
        return lambda *a, **kw: self._with(name, *a, **kw)

    def __repr__(self):
        return f"<modal.Image {len(self._steps)} recorded steps>"
