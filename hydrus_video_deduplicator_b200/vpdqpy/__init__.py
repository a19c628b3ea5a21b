from .vpdqpy import DOWNSCALE_DIMENSIONS, Vpdq, VpdqHash  # noqa: F401
