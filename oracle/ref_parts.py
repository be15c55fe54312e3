"""ctypes wrapper of oracle/_ref/libdsopp_ref_parts.so -- the two pieces of the reference itself that compile here
(oracle/build_ref.py).  Test infrastructure only."""
import ctypes as C

import numpy as np

from . import build_ref

CALL_NAMES = ["energy", "linearize", "step", "accept", "reject"]
_lib = None


def available():
    return build_ref.available()


def load():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("neither /root/reference nor a prebuilt oracle/_ref library is present")
        lib = C.CDLL(path)
        lib.ref_pixelinfo_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_pixelinfo_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_lm_solve.restype = C.c_int
        lib.ref_lm_solve.argtypes = ([C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double]
                                     + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
                                     + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
        _lib = lib
    return _lib


def pattern():
    """dsopp::Pattern -> ((8, 2) offsets (x_i, y_i), centre index)."""
    lib = load()
    xy = np.zeros(16)
    c = C.c_int()
    lib.ref_pattern.restype = C.c_int
    lib.ref_pattern.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    n = lib.ref_pattern(xy.ctypes.data, C.byref(c))
    return xy.reshape(n, 2), c.value


def _aligned(shape, dtype, align=32, offset=0):
    """Array whose data pointer is `offset` bytes past a multiple of `align` (the reference picks its AVX2 path by
    alignment, calculate_pixelinfo.cpp:386-392)."""
    n = int(np.prod(shape))
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + align + offset, np.uint8)
    start = (-raw.ctypes.data) % align + offset
    return raw[start:start + n * item].view(dtype).reshape(shape)


def pixelinfo(image, aligned=True):
    """dsopp::features::calculate_pixelinfo<1>: (H, W) -> (H, W, 3) interleaved {I, dx, dy}, float64 or float32."""
    lib = load()
    img = np.asarray(image)
    assert img.dtype in (np.float64, np.float32) and img.ndim == 2
    H, W = img.shape
    # Reference quirk: the dispatch `width % 8 == 0 && is_aligned(input, 32), is_aligned(output, 32)`
    # (calculate_pixelinfo.cpp:388) is a COMMA expression -- the alignment of `output` alone selects the AVX2 kernel, which
    # then runs past the rows when width % 8 != 0.  The reference only ever feeds it pyramid widths that are multiples
    # of 8; do the same here instead of crashing.
    if aligned and img.dtype == np.float64 and W % 8 != 0:
        raise ValueError("the reference's AVX2 path needs width % 8 == 0 (use aligned=False for the plain-C path)")
    off = 0 if aligned else img.dtype.itemsize
    src = _aligned((H, W), img.dtype, offset=off)
    src[...] = img
    dst = _aligned((H, W, 3), img.dtype, offset=off)
    fn = lib.ref_pixelinfo_f64 if img.dtype == np.float64 else lib.ref_pixelinfo_f32
    fn(src.ctypes.data, dst.ctypes.data, W, H)
    return dst.copy()


def lm_solve(energies, valid, norms, max_it=50, lambda0=1e-5, ftol=1e-8, ptol=1e-8, force_accept=False, min_it=0,
             dec=2.0, inc=10.0):
    """levenberg_marquardt_algorithm::solve on a scripted problem -> (calls, lambdas, energy, n_valid, converged)."""
    lib = load()
    e = np.ascontiguousarray(energies, np.float64)
    v = np.ascontiguousarray(valid, np.int32)
    nr = np.ascontiguousarray(norms, np.float64).reshape(-1, 2)
    calls = np.zeros(8 * (max_it + 2), np.int32)
    lams = np.zeros(max_it + 2, np.float64)
    out_e, out_v, out_c = C.c_double(), C.c_int32(), C.c_int32()
    n = lib.ref_lm_solve(max_it, lambda0, ftol, ptol, int(force_accept), min_it, dec, inc, e.ctypes.data, v.ctypes.data,
                         len(e), nr.ctypes.data, len(nr), calls.ctypes.data, len(calls), lams.ctypes.data, len(lams),
                         C.byref(out_e), C.byref(out_v), C.byref(out_c))
    assert n <= len(calls)
    calls = calls[:n]
    return [CALL_NAMES[c] for c in calls], lams[:int((calls == 2).sum())].copy(), out_e.value, out_v.value, bool(out_c.value)
