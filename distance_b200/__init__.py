"""distance_b200 -- B200 (sm_100a) engine for the pairwise-comparison hot path of
benjamincjackson/distance.

The product is the C-ABI shared library built from distance_b200/csrc (see include/distance_gpu.h)
and the C++ `distance` CLI host on top of it.  This Python package is a thin ctypes view of the
same C ABI, used by the parity tests and bench.py; it holds no compute of its own and never falls
back to the CPU: if the library is missing it raises.
"""
from .api import (  # noqa: F401
    DG_INPUT_ASCII,
    DG_INPUT_NIBBLE,
    DG_INPUT_PARADIS,
    MEASURES,
    DistanceGpuError,
    Engine,
    build_library,
    device_count,
    library_path,
    load_library,
    pack_nibbles,
)

__all__ = [
    "Engine", "DistanceGpuError", "MEASURES", "DG_INPUT_ASCII", "DG_INPUT_PARADIS", "DG_INPUT_NIBBLE", "pack_nibbles",
    "build_library", "load_library", "library_path", "device_count",
]
