"""Web decorators: OUT OF SCOPE for the hot path -- accepted and inert so
that scripts which also define endpoints still import; the decorated callable stays callable."""
from .cls import _mark


def _inert(kind):
    def outer(*dargs, **dkw):
        if len(dargs) == 1 and callable(dargs[0]) and not dkw:
            return _mark(dargs[0], web=kind)
        return lambda fn: _mark(fn, web=kind, web_kwargs=dkw)

    return outer


asgi_app = _inert("asgi_app")
wsgi_app = _inert("wsgi_app")
fastapi_endpoint = _inert("fastapi_endpoint")
web_endpoint = _inert("web_endpoint")
web_server = _inert("web_server")
