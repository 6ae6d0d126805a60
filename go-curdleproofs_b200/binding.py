"""ctypes mirror of include/curdle_b200.h.  Buffers are raw bytes in gnark's
memory layout (see the header); no arithmetic happens on this side."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "curdle_b200.h")

FP_BYTES = 48
FR_BYTES = 32
G1_AFFINE_BYTES = 96
G1_JAC_BYTES = 144


class CdlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"curdle_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


def lib_path() -> str:
    return os.path.join(HERE, "libcurdle_b200.so")


def declared_symbols():
    """Every function the public header declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cdl_[a-z0-9_]+)\s*\(", txt)))


_LIB = None


def load_library():
    """dlopen libcurdle_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(p)
    vp, sz, i32, u32p = C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_uint32)
    lib.cdl_abi_version.restype = C.c_uint32
    lib.cdl_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.cdl_destroy.argtypes = [vp]
    lib.cdl_destroy.restype = None
    lib.cdl_last_error.argtypes = [vp]
    lib.cdl_last_error.restype = C.c_char_p
    lib.cdl_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.c_char_p, sz]
    lib.cdl_g1_msm.argtypes = [vp, vp, vp, sz, vp]
    lib.cdl_g1_msm_batch.argtypes = [vp, vp, vp, u32p, sz, vp]
    lib.cdl_g1_scalar_mul_affine.argtypes = [vp, vp, vp, sz, sz, vp]
    lib.cdl_g1_fold.argtypes = [vp, vp, vp, vp, sz]
    lib.cdl_g1_batch_to_affine.argtypes = [vp, vp, sz, vp]
    lib.cdl_g1_sum_affine.argtypes = [vp, vp, sz, vp]
    lib.cdl_g1_compress.argtypes = [vp, vp, sz, vp]
    lib.cdl_g1_decompress.argtypes = [vp, vp, sz, vp, vp]
    lib.cdl_fp_mul.argtypes = [vp, vp, vp, sz, vp]
    lib.cdl_int_peak.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in declared_symbols():
        fn = getattr(lib, name)  # AttributeError if the build lost a symbol
        if name not in ("cdl_destroy", "cdl_last_error", "cdl_abi_version"):
            fn.restype = i32
    _LIB = lib
    return lib


class Context:
    """One CUDA device; mirrors cdl_ctx."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.cdl_create(device, C.byref(h))
        if rc != 0:
            raise CdlError(rc, "cdl_create failed (no CUDA device? this library has no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.cdl_destroy(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc: int):
        if rc != 0:
            raise CdlError(rc, (self.lib.cdl_last_error(self.h) or b"").decode())

    def device_info(self):
        sm, khz = C.c_int32(), C.c_int32()
        name = C.create_string_buffer(128)
        self._chk(self.lib.cdl_device_info(self.h, C.byref(sm), C.byref(khz), name, 128))
        return {"sm_count": sm.value, "clock_khz": khz.value, "name": name.value.decode()}

    # --- 1:1 gnark replacements; all arguments/results are bytes -------------
    def g1_msm(self, points: bytes, scalars: bytes) -> bytes:
        n = len(points) // G1_AFFINE_BYTES
        if len(scalars) != n * FR_BYTES:
            raise CdlError(-3, "len(points) != len(scalars)")
        out = C.create_string_buffer(G1_JAC_BYTES)
        self._chk(self.lib.cdl_g1_msm(self.h, points, scalars, n, out))
        return out.raw

    def g1_msm_batch(self, points: bytes, scalars: bytes, offsets) -> bytes:
        k = len(offsets) - 1
        offs = (C.c_uint32 * (k + 1))(*offsets)
        out = C.create_string_buffer(max(1, k) * G1_AFFINE_BYTES)
        self._chk(self.lib.cdl_g1_msm_batch(self.h, points, scalars, offs, k, out))
        return out.raw[: k * G1_AFFINE_BYTES]

    def g1_scalar_mul_affine(self, points: bytes, scalars: bytes, broadcast: bool) -> bytes:
        n = len(points) // G1_AFFINE_BYTES
        out = C.create_string_buffer(max(1, n) * G1_AFFINE_BYTES)
        self._chk(self.lib.cdl_g1_scalar_mul_affine(self.h, points, scalars, n, 0 if broadcast else 1, out))
        return out.raw[: n * G1_AFFINE_BYTES]

    def g1_fold(self, L: bytes, R: bytes, x: bytes) -> bytes:
        n = len(L) // G1_AFFINE_BYTES
        buf = C.create_string_buffer(L, max(1, len(L)))
        self._chk(self.lib.cdl_g1_fold(self.h, buf, R, x, n))
        return buf.raw[: len(L)]

    def g1_batch_to_affine(self, jac: bytes) -> bytes:
        n = len(jac) // G1_JAC_BYTES
        out = C.create_string_buffer(max(1, n) * G1_AFFINE_BYTES)
        self._chk(self.lib.cdl_g1_batch_to_affine(self.h, jac, n, out))
        return out.raw[: n * G1_AFFINE_BYTES]

    def g1_sum_affine(self, points: bytes) -> bytes:
        n = len(points) // G1_AFFINE_BYTES
        out = C.create_string_buffer(G1_AFFINE_BYTES)
        self._chk(self.lib.cdl_g1_sum_affine(self.h, points, n, out))
        return out.raw

    def g1_compress(self, points: bytes) -> bytes:
        n = len(points) // G1_AFFINE_BYTES
        out = C.create_string_buffer(max(1, n) * 48)
        self._chk(self.lib.cdl_g1_compress(self.h, points, n, out))
        return out.raw[: n * 48]

    def g1_decompress(self, enc: bytes, check: bool = True):
        """Returns (affine bytes, status list).  Raises CdlError(CDL_ERR_DECODE)
        when `check` and any point was rejected."""
        n = len(enc) // 48
        out = C.create_string_buffer(max(1, n) * G1_AFFINE_BYTES)
        st = C.create_string_buffer(max(1, n))
        rc = self.lib.cdl_g1_decompress(self.h, enc, n, out, st)
        if rc != 0 and (check or rc != -4):
            self._chk(rc)
        return out.raw[: n * G1_AFFINE_BYTES], list(st.raw[:n])

    def fp_mul(self, a: bytes, b: bytes) -> bytes:
        n = len(a) // FP_BYTES
        out = C.create_string_buffer(max(1, n) * FP_BYTES)
        self._chk(self.lib.cdl_fp_mul(self.h, a, b, n, out))
        return out.raw[: n * FP_BYTES]

    def int_peak(self, kind: int, iters: int):
        ops, ms = C.c_double(), C.c_double()
        self._chk(self.lib.cdl_int_peak(self.h, kind, iters, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value
