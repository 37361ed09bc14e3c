from . import transform  # noqa: F401
