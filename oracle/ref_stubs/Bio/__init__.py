from . import motifs  # noqa: F401
