from . import lightning  # noqa: F401
