"""richmol_b200 -- B200-native (sm_100a) implementation of richmol's field-driven TDSE hot path.

Drop-in for `richmol.field.CarTens` (field / vec / mul / sums / tomat) and `richmol.tdse.TDSE`
(time_grid / init_state / update); everything on the per-step path runs in hand-written CUDA behind
the C ABI of `include/richmol_b200.h` (librichmol_b200.so).  There is no CPU fallback: using the
hot path without the built library or without a CUDA device raises.
"""
from . import convert_units  # noqa: F401
from .field import CarTens  # noqa: F401
from .tdse import TDSE  # noqa: F401

__version__ = "0.1.0"
