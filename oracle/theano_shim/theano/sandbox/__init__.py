from . import rng_mrg  # noqa: F401
