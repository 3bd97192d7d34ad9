from . import measure  # noqa: F401
