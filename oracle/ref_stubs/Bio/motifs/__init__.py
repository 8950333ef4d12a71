from . import jaspar  # noqa: F401
