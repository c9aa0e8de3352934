"""isochrones_b200 — B200-native implementation of the lnpost hot path of timothydmorton/isochrones.

Python host code (this package, mirroring the reference's call signatures for the path) over a ctypes C ABI
(``include/isochrones_b200.h``) over hand-written sm_100a CUDA kernels (``csrc/``).  No PyTorch, no Triton, no
CPU fallback: the compute entry points raise when ``lib/libisochrones_b200.so`` or a CUDA device is missing.
"""
__version__ = "0.1.0"

from .bcio import load_mist_bc_grid  # noqa: F401
from .interp import DFInterpolator  # noqa: F401
from .models import (EvolutionTrackInterpolator, IsochroneInterpolator, ModelGridInterpolator,  # noqa: F401
                     get_ichrone, ichrone_from_arrays)
from .starmodel import (BasicStarModel, BinaryStarModel, SingleStarModel, TripleStarModel,  # noqa: F401
                        compile_catalog)
