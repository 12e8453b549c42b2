"""draco_sharp_b200: B200-native (sm_100a) batch implementation of draco-sharp's attribute-decode hot path.

Product code lives in csrc/ (CUDA kernels + the C ABI of include/dracob200.h); this package is the
thin host-side mirror of the reference interface over ctypes.  No CPU decode path exists here.
"""
from ._native import DracoError, lib  # noqa: F401
from .decoder import Batch, Draco, DracoBatchDecoder, DracoHeader, PointAttribute, index_only  # noqa: F401

__all__ = ["DracoBatchDecoder", "Draco", "DracoHeader", "PointAttribute", "Batch", "DracoError", "index_only", "lib"]
