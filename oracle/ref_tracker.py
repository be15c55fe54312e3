"""ctypes face of oracle/_ref/libdsopp_ref_tracker.so (oracle/build_ref_tracker.py): the reference's own
createReferenceDepthMaps and optimizeImmatureLandmark.  Test infrastructure only."""
import ctypes as C

import numpy as np

from . import build_ref_tracker

_lib = None


def available():
    return build_ref_tracker.available()


def load():
    global _lib
    if _lib is None:
        path = build_ref_tracker.build()
        if path is None:
            raise RuntimeError("neither /root/reference nor a prebuilt oracle/_ref/libdsopp_ref_tracker.so is present")
        _lib = C.CDLL(path)
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def create_reference_depth_maps(T_w_agent, intr, width, height, levels, landmarks):
    """T_w_agent: n poses (4x4 or 3x4), the last one is the target.  landmarks: for each of the n - 1 older keyframes a dict
    of uv (M, 2), idepth, idepth_variance, outlier, marginalized, status (the connection to the target), as the TRACK holds
    them.  -> list over levels of (idepth_w, weight), each (H >> l, W >> l)."""
    lib = load()
    n = len(T_w_agent)
    T = _f64([np.asarray(t)[:3, :4] for t in T_w_agent])
    off = np.zeros(n, dtype=np.int32)
    off[1:] = np.cumsum([len(l["idepth"]) for l in landmarks])
    cat = lambda k, f: f(np.concatenate([np.asarray(l[k]).reshape(len(l["idepth"]), -1) for l in landmarks]))  # noqa: E731
    uv, idp, var = cat("uv", _f64), cat("idepth", _f64), cat("idepth_variance", _f64)
    outl, marg, st = cat("outlier", _u8), cat("marginalized", _u8), cat("status", _u8)
    sizes = [(height >> l, width >> l) for l in range(levels)]
    out = np.zeros(2 * sum(h * w for h, w in sizes))
    k = _f64(intr)
    vp, i = C.c_void_p, C.c_int
    lib.reftrk_create_reference_depth_maps.restype = i
    lib.reftrk_create_reference_depth_maps.argtypes = [i, vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    got = lib.reftrk_create_reference_depth_maps(n, T.ctypes.data, k.ctypes.data, width, height, levels, off.ctypes.data,
                                                 uv.ctypes.data, idp.ctypes.data, var.ctypes.data, outl.ctypes.data,
                                                 marg.ctypes.data, st.ctypes.data, out.ctypes.data)
    assert got == levels
    res, o = [], 0
    for h, w in sizes:
        res.append((out[o:o + h * w].reshape(h, w).copy(), out[o + h * w:o + 2 * h * w].reshape(h, w).copy()))
        o += 2 * h * w
    return res


def optimize_immature_landmark(T_w_agent, exposure, affine, images, masks, intr, ref_index, projection, patch, idepth_min,
                               idepth_max, minimum_inliers, sigma_huber):
    """images: (n, H, W) raw level-0 intensities; masks: (n, H, W) uint8 or None.
    -> (status: 0 activate / 2 delete, idepth afterwards)"""
    lib = load()
    n = len(T_w_agent)
    T = _f64([np.asarray(t)[:3, :4] for t in T_w_agent])
    im = _f64(images)
    _, H, W = im.shape
    m = None if masks is None else _u8(masks)
    e, ab, k, pr, pt = _f64(exposure), _f64(affine), _f64(intr), _f64(projection), _f64(patch)
    out = C.c_double()
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.reftrk_optimize_immature_landmark.restype = i
    lib.reftrk_optimize_immature_landmark.argtypes = [i, vp, vp, vp, vp, vp, vp, i, i, i, vp, vp, d, d, i, d, vp]
    st = lib.reftrk_optimize_immature_landmark(n, T.ctypes.data, e.ctypes.data, ab.ctypes.data, im.ctypes.data,
                                               None if m is None else m.ctypes.data, k.ctypes.data, W, H, int(ref_index),
                                               pr.ctypes.data, pt.ctypes.data, float(idepth_min), float(idepth_max),
                                               int(minimum_inliers), float(sigma_huber), C.addressof(out))
    return st, out.value
