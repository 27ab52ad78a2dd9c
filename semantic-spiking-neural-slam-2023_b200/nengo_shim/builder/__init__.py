from . import ensemble  # noqa: F401
