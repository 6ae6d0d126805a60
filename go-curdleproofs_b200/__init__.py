"""go-curdleproofs_b200 — B200-native BLS12-381 G1 hot path for go-curdleproofs.

Python here is only a thin ctypes mirror of the C ABI in include/curdle_b200.h
(the product is libcurdle_b200.so: hand-written CUDA for sm_100a + the C++ host
engine).  There is no CPU fallback: loading fails loudly when the library has
not been built, and creating a context fails when no CUDA device is present.
"""
from __future__ import annotations

from .binding import (  # noqa: F401
    CdlError,
    Context,
    CRS,
    Rand,
    DeviceBuffer,
    comm_unique_id,
    comm_partition,
    G1_AFFINE_BYTES,
    G1_JAC_BYTES,
    FR_BYTES,
    FP_BYTES,
    lib_path,
    load_library,
    declared_symbols,
)
