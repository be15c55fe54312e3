"""ctypes binding of include/dsopp_cuda_pose_alignment.h (test / bench plumbing; no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_P, _I, _D = C.c_void_p, C.c_int32, C.c_double


class Config(C.Structure):
    _fields_ = [("max_points", _I), ("max_width", _I), ("max_height", _I), ("device", _I)]


class Options(C.Structure):
    _fields_ = [("max_num_iterations", _I), ("initial_trust_region_radius", _D), ("function_tolerance", _D),
                ("parameter_tolerance", _D), ("sigma_huber_loss", _D), ("affine_brightness_regularizer", _D * 2),
                ("regularizer_decrease_on_accept", _D), ("regularizer_increase_on_reject", _D)]


class Result(C.Structure):
    _fields_ = [("rmse", _D), ("energy", _D), ("number_of_valid_residuals", _I), ("converged", _I), ("iterations", _I),
                ("T_target_reference", _D * 12), ("T_world_target", _D * 12), ("affine_brightness_eps", _D * 2),
                ("hessian", _D * 64)]


# name -> (restype, argtypes); exactly the symbols declared in include/dsopp_cuda_pose_alignment.h
SIGNATURES = {
    "dpa_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "dpa_destroy": (C.c_int, [_P]),
    "dpa_last_error": (C.c_char_p, [_P]),
    "dpa_stream": (_P, [_P]),
    "dpa_set_reference_landmarks": (C.c_int, [_P, _I, _P, _P, _P, _P, _D, _P, _P, _I, _I]),
    "dpa_set_reference_depth_map": (C.c_int, [_P, _P, _P, _P, _P, _D, _P, _P, _I, _I]),
    "dpa_num_landmarks": (C.c_int, [_P]),
    "dpa_get_reference_landmarks": (C.c_int, [_P, _I, _P, _P, _P]),
    "dpa_set_target": (C.c_int, [_P, _P, _P, _P, _D, _P, _P, _I, _I]),
    "dpa_solve": (C.c_int, [_P, C.POINTER(Options), _P, C.POINTER(Result)]),
    "dpa_get_trace": (C.c_int, [_P, _I, _P, _P, _P]),
    "dpa_mean_square_optical_flow": (C.c_int, [_P, _P, C.POINTER(_D), C.POINTER(C.c_int32)]),
    "dpa_set_grid_threshold": (C.c_int, [_P, _I]),
}

_bound = False


def _lib():
    global _bound
    lib = capi.load_library()
    if not _bound:
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return lib


def default_options(sigma=20.0, ab_reg=(1e12, 1e8), max_it=50):
    """createPoseAlignment, tracker/tracker/src/fabric.cpp:127-147 + EigenPoseAlignment::solve :298-305."""
    return Options(max_it, 1e2, 1e-5, 1e-5, sigma, (_D * 2)(*ab_reg), 2.0, 2.0)


class Aligner:
    def __init__(self, max_points, max_width, max_height, device=0):
        self.lib = _lib()
        self.h = C.c_void_p()
        cfg = Config(max_points, max_width, max_height, device)
        rc = self.lib.dpa_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise capi.DpbaError(f"dpa_create failed with {rc} (is a CUDA device visible? there is no CPU fallback)")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.dpa_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise capi.DpbaError(f"dpa error {rc}: {self.lib.dpa_last_error(self.h).decode()}")
        return rc

    @property
    def stream(self):
        return self.lib.dpa_stream(self.h)

    def set_reference_landmarks(self, xy, idepth, patch, T_w, exposure, ab0, intr, width, height):
        xy, idepth, patch = capi._f32(xy), capi._f32(idepth), capi._f32(patch)
        T, ab, it = capi.pose34(T_w), capi._f64(ab0), capi._f64(intr)
        self._ck(self.lib.dpa_set_reference_landmarks(self.h, len(idepth), capi._ptr(xy), capi._ptr(idepth), capi._ptr(patch),
                                                      capi._ptr(T), exposure, capi._ptr(ab), capi._ptr(it), width, height))

    def set_reference_depth_map(self, image, idepth_sum, weight, T_w, exposure, ab0, intr):
        image, ids, w = capi._f32(image), capi._f32(idepth_sum), capi._f32(weight)
        T, ab, it = capi.pose34(T_w), capi._f64(ab0), capi._f64(intr)
        H, W = w.shape
        return self._ck(self.lib.dpa_set_reference_depth_map(self.h, capi._ptr(image), capi._ptr(ids), capi._ptr(w), capi._ptr(T),
                                                             exposure, capi._ptr(ab), capi._ptr(it), W, H))

    def num_landmarks(self):
        return self._ck(self.lib.dpa_num_landmarks(self.h))

    def get_reference_landmarks(self):
        n = self.num_landmarks()
        xy, idepth, patch = np.zeros((n, 2), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        self._ck(self.lib.dpa_get_reference_landmarks(self.h, n, capi._ptr(xy), capi._ptr(idepth), capi._ptr(patch)))
        return xy, idepth, patch

    def set_target(self, image, mask, T_w, exposure, ab0, intr):
        image, mask = capi._f32(image), capi._u8(mask)
        T, ab, it = capi.pose34(T_w), capi._f64(ab0), capi._f64(intr)
        H, W = image.shape[:2]
        self._ck(self.lib.dpa_set_target(self.h, capi._ptr(image), capi._ptr(mask), capi._ptr(T), exposure, capi._ptr(ab),
                                         capi._ptr(it), W, H))

    def set_grid_threshold(self, min_points):
        self._ck(self.lib.dpa_set_grid_threshold(self.h, int(min_points)))

    def mean_square_optical_flow(self, T_target_reference):
        """calculateMeanSquareOpticalFlow over the resident reference landmarks -> (flow, landmarks used)."""
        T = capi.pose34(T_target_reference)
        flow, n = _D(), C.c_int32()
        self._ck(self.lib.dpa_mean_square_optical_flow(self.h, capi._ptr(T), C.byref(flow), C.byref(n)))
        return flow.value, n.value

    def trace(self):
        e, lam, acc = np.zeros(64), np.zeros(64), np.zeros(64, np.int32)
        n = self._ck(self.lib.dpa_get_trace(self.h, 64, capi._ptr(e), capi._ptr(lam), capi._ptr(acc)))
        return [dict(energy=float(e[i]), lam=float(lam[i]), accepted=bool(acc[i])) for i in range(n)]

    def solve(self, options=None, prior_rotation=None):
        o = options or default_options()
        r = Result()
        pr = capi._f64(prior_rotation)
        self._ck(self.lib.dpa_solve(self.h, C.byref(o), capi._ptr(pr), C.byref(r)))
        to44 = lambda a: np.vstack([np.array(a[:]).reshape(3, 4), [0, 0, 0, 1.0]])  # noqa: E731
        return dict(rmse=r.rmse, energy=r.energy, n_valid=r.number_of_valid_residuals, converged=bool(r.converged),
                    iterations=r.iterations, T_t_r=to44(r.T_target_reference), T_w_target=to44(r.T_world_target),
                    ab_eps=np.array(r.affine_brightness_eps[:]), H=np.array(r.hessian[:]).reshape(8, 8))
