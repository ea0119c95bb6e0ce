"""``modal.exception`` names the reference scripts catch."""


class Error(Exception):
    pass


class NotFoundError(Error):
    pass


class InvalidError(Error):
    pass


class ExecutionError(Error):
    pass


class FunctionTimeoutError(Error, TimeoutError):
    pass


class InputCancellation(BaseException):
    pass


class DeserializationError(Error):
    pass


class RemoteError(Error):
    pass


class AuthError(Error):
    pass


class ConnectionError(Error):  # noqa: A001
    pass


class TimeoutError(Error):  # noqa: A001
    pass
