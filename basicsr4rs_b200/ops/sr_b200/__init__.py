from . import raw  # noqa: F401
