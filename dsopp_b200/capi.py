"""ctypes binding of the C ABI in include/dsopp_cuda_pba.h (test / bench plumbing).

The product path is the CUDA library: if it is missing or no GPU is present this module raises --
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(HERE, "lib", "libdsopp_pba_cuda.so")
HOST_LIB_PATH = os.path.join(HERE, "lib", "libdsopp_pba_host.so")

MAX_FRAMES = 16
BLOCK = 8


class DpbaError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("max_frames", C.c_int32),
        ("max_points_per_frame", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
    ]


class ResidualView(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("residuals", C.c_void_p),
        ("d_reference_state_eps", C.c_void_p),
        ("d_target_state_eps", C.c_void_p),
        ("d_idepth", C.c_void_p),
        ("huber_weight", C.c_void_p),
        ("energy", C.c_void_p),
        ("connection_status", C.c_void_p),
        ("connection_status_candidate", C.c_void_p),
    ]


_P = C.c_void_p
_I = C.c_int32
_D = C.c_double


class LmOptions(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32),
        ("min_num_iterations", C.c_int32),
        ("force_accept", C.c_int32),
        ("first_estimate_jacobians", C.c_int32),
        ("initial_levenberg_marquardt_regularizer", C.c_double),
        ("function_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("levenberg_marquardt_regularizer_decrease_on_accept", C.c_double),
        ("levenberg_marquardt_regularizer_increase_on_reject", C.c_double),
        ("sigma_huber_loss", C.c_double),
        ("affine_brightness_regularizer", C.c_double * 2),
        ("fixed_state_regularizer", C.c_double),
    ]


class LmResult(C.Structure):
    _fields_ = [("energy", C.c_double), ("number_of_valid_residuals", C.c_int32), ("converged", C.c_int32),
                ("iterations", C.c_int32)]

# name -> (restype, argtypes); exactly the symbols declared in include/dsopp_cuda_pba.h
SIGNATURES = {
    "dpba_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "dpba_destroy": (C.c_int, [_P]),
    "dpba_last_error": (C.c_char_p, [_P]),
    "dpba_version": (C.c_char_p, []),
    "dpba_stream": (_P, [_P]),
    "dpba_push_frame": (C.c_int, [_P, _I, _P, _P, _P, _D, _P, _P, _I]),
    "dpba_push_frame_intensity": (C.c_int, [_P, _I, _P, _P, _P, _D, _P, _P, _I]),
    "dpba_push_frame_raw": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _D, _P, _P, _I]),
    "dpba_build_pyramid": (C.c_int, [_P, _P, _P, _P, _I, _P]),
    "dpba_remove_frame": (C.c_int, [_P, _I]),
    "dpba_num_frames": (C.c_int, [_P]),
    "dpba_synchronize": (C.c_int, [_P]),
    "dpba_set_frame_linearization": (C.c_int, [_P, _I, _P, _P]),
    "dpba_set_frame_flags": (C.c_int, [_P, _I, _I, _I]),
    "dpba_set_frame_marginalized": (C.c_int, [_P, _I, _I]),
    "dpba_set_landmarks": (C.c_int, [_P, _I, _I, _P, _P, _P, _P]),
    "dpba_append_landmarks": (C.c_int, [_P, _I, _I, _P, _P, _P, _P]),
    "dpba_set_landmark_flags": (C.c_int, [_P, _I, _I, _P]),
    "dpba_num_landmarks": (C.c_int, [_P, _I]),
    "dpba_get_landmarks": (C.c_int, [_P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "dpba_get_pose_idepth_blocks": (C.c_int, [_P, _I, _I, _P]),
    "dpba_set_statuses": (C.c_int, [_P, _I, _I, _I, _P]),
    "dpba_get_statuses": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "dpba_append_statuses": (C.c_int, [_P, _I, _I, _I, _I, _P]),
    "dpba_get_residual_scalars": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "dpba_set_frame_statuses": (C.c_int, [_P, _I, _I, _P]),
    "dpba_set_window_landmarks": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "dpba_get_frame_statuses": (C.c_int, [_P, _I, _I, _P, _P]),
    "dpba_set_state": (C.c_int, [_P, _P, _P]),
    "dpba_get_state": (C.c_int, [_P, _P, _P]),
    "dpba_first_estimate": (C.c_int, [_P]),
    "dpba_evaluate": (C.c_int, [_P, _D, _I, _I, C.POINTER(_D), C.POINTER(_I)]),
    "dpba_evaluate_jacobians": (C.c_int, [_P, _D, _I, _I]),
    "dpba_download_residual_block": (C.c_int, [_P, _I, _I, C.POINTER(ResidualView)]),
    "dpba_linearize": (C.c_int, [_P, _D, _I, _I, _I, _P, _P, _P, _P]),
    "dpba_linearize_materialized": (C.c_int, [_P, _D, _I, _I, _I, _P, _P, _P, _P]),
    "dpba_back_substitute": (C.c_int, [_P, _P, _D]),
    "dpba_accept": (C.c_int, [_P, C.POINTER(_D), C.POINTER(_D)]),
    "dpba_reject": (C.c_int, [_P]),
    "dpba_change_residual_statuses": (C.c_int, [_P, _I]),
    "dpba_landmarks_energy": (C.c_int, [_P, _I, C.POINTER(_D), C.POINTER(_I)]),
    "dpba_update_point_statuses": (C.c_int, [_P, _I, _D, C.POINTER(_D)]),
    "dpba_refine_immature_landmarks": (C.c_int, [_P, _I, _I, _P, _P, _P, _I, _D, _P, _P, _P]),
    "dpba_solve_lm": (C.c_int, [_P, C.POINTER(LmOptions), _P, _P, _D, C.POINTER(LmResult)]),
    "dpba_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "dpba_launch_count": (C.c_int64, []),
    "dpba_debug_stamps": (C.c_int, [_I, _P]),
    "dpba_debug_cta_times": (C.c_int, [_P, _I]),
    "dpba_debug_kernel_times": (C.c_int, [_P]),
    "dpba_debug_pixelinfo_ab": (C.c_int, [_I, _I, _I, _P, _P]),
    "dpba_profile_enable": (C.c_int, [_P, _I]),
    "dpba_profile_read": (C.c_int, [_P, _P, _P]),
    "dpba_comm_unique_id": (C.c_int, [_P]),
    "dpba_comm_init": (C.c_int, [_P, _P, _I, _I]),
    "dpba_create_reference_depth_maps": (C.c_int, [_P, _I, C.c_double, _P, _P]),
    "dpba_peer_export": (C.c_int, [_P, _P]),
    "dpba_peer_barrier": (C.c_int, [_P]),
    "dpba_peer_attach": (C.c_int, [_P, _P, _I, _I]),
}

_lib = None


def load_library(path: str = CUDA_LIB_PATH):
    """dlopen the CUDA library and bind every declared entry point; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise DpbaError(f"{path} is not built: run `python -m dsopp_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    """Address of a NumPy array for a c_void_p parameter (None -> NULL).  The plain integer is about twice as cheap as
    ndarray.ctypes.data_as(), and an upload / readback makes ~130 of these.  The caller keeps `a` referenced until the C
    call has returned (every wrapper below binds its converted arrays to local names first)."""
    return None if a is None else a.ctypes.data


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def pose34(T):
    """4x4 (or 3x4) pose -> 12 doubles, 3x4 row-major."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:3, :4]).reshape(12)


class Handle:
    """Thin RAII wrapper: one dpba_handle.  Method names follow the C entry points."""

    def __init__(self, max_frames, max_points_per_frame, width, height, device=0, rank=0, world_size=1):
        self.lib = load_library()
        self.cfg = Config(max_frames, max_points_per_frame, width, height, device, rank, world_size)
        self.h = C.c_void_p()
        rc = self.lib.dpba_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise DpbaError(f"dpba_create failed with {rc} (is a CUDA device visible? there is no CPU fallback)")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.dpba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise DpbaError(f"dpba error {rc}: {self.lib.dpba_last_error(self.h).decode()}")
        return rc

    @property
    def n_frames(self):
        return self.lib.dpba_num_frames(self.h)

    @property
    def stream(self):
        return self.lib.dpba_stream(self.h)

    def push_frame(self, frame_id, image, mask, T_w_lin, exposure, ab0, intr, fixed):
        image = _f32(image)
        mask = _u8(mask)
        T, ab, it = pose34(T_w_lin), _f64(ab0), _f64(intr)
        if image.ndim == 3:
            return self._ck(self.lib.dpba_push_frame(self.h, frame_id, _ptr(image), _ptr(mask), _ptr(T), exposure,
                                                     _ptr(ab), _ptr(it), int(fixed)))
        return self._ck(self.lib.dpba_push_frame_intensity(self.h, frame_id, _ptr(image), _ptr(mask), _ptr(T),
                                                           exposure, _ptr(ab), _ptr(it), int(fixed)))

    def push_frame_raw(self, frame_id, gray, lut, vignetting, mask, T_w_lin, exposure, ab0, intr, fixed):
        gray, vignetting, mask, lut = _u8(gray), _u8(vignetting), _u8(mask), _f32(lut)
        T, ab, it = pose34(T_w_lin), _f64(ab0), _f64(intr)
        return self._ck(self.lib.dpba_push_frame_raw(self.h, frame_id, _ptr(gray), _ptr(lut), _ptr(vignetting), _ptr(mask),
                                                     _ptr(T), exposure, _ptr(ab), _ptr(it), int(fixed)))

    def build_pyramid(self, gray, lut=None, vignetting=None, levels=4):
        gray, vignetting, lut = _u8(gray), _u8(vignetting), _f32(lut)
        H, W = gray.shape
        levels = min(levels, 5)
        outs = [np.zeros((H >> l, W >> l, 3), np.float32) for l in range(levels)]
        ptrs = (C.c_void_p * levels)(*[o.ctypes.data for o in outs])
        self._ck(self.lib.dpba_build_pyramid(self.h, _ptr(gray), _ptr(lut), _ptr(vignetting), levels, C.cast(ptrs, C.c_void_p)))
        return outs

    def synchronize(self):
        self._ck(self.lib.dpba_synchronize(self.h))

    def remove_frame(self, slot):
        self._ck(self.lib.dpba_remove_frame(self.h, slot))

    def set_frame_linearization(self, slot, T_w_lin, ab0):
        T, ab = pose34(T_w_lin), _f64(ab0)
        self._ck(self.lib.dpba_set_frame_linearization(self.h, slot, _ptr(T), _ptr(ab)))

    def set_frame_flags(self, slot, fixed, to_marginalize):
        self._ck(self.lib.dpba_set_frame_flags(self.h, slot, int(fixed), int(to_marginalize)))

    def set_landmarks(self, slot, uv, idepth, patch, flags=None, append=False):
        uv, idepth, patch, flags = _f32(uv), _f32(idepth), _f32(patch), _u8(flags)
        fn = self.lib.dpba_append_landmarks if append else self.lib.dpba_set_landmarks
        self._ck(fn(self.h, slot, len(idepth), _ptr(uv), _ptr(idepth), _ptr(patch), _ptr(flags)))

    def set_landmark_flags(self, slot, flags):
        flags = _u8(flags)
        self._ck(self.lib.dpba_set_landmark_flags(self.h, slot, len(flags), _ptr(flags)))

    def num_landmarks(self, slot):
        return self._ck(self.lib.dpba_num_landmarks(self.h, slot))

    def get_landmarks(self, slot):
        n = self.num_landmarks(slot)
        f32 = np.zeros((5, n), np.float32)  # one block, five contiguous rows: one address lookup instead of five
        flags, n_inl = np.zeros(n, np.uint8), np.zeros(n, np.uint32)
        b, row = f32.ctypes.data, 4 * n
        self._ck(self.lib.dpba_get_landmarks(self.h, slot, n, b, b + row, b + 2 * row, b + 3 * row, _ptr(flags),
                                             _ptr(n_inl), b + 4 * row))
        return dict(idepth=f32[0], idepth_step=f32[1], inv_hdd=f32[2], b_d=f32[3], flags=flags, n_inliers=n_inl,
                    rel_baseline=f32[4])

    def get_pose_idepth_blocks(self, slot):
        n = self.num_landmarks(slot)
        out = np.zeros((n, BLOCK * self.n_frames), np.float32)
        self._ck(self.lib.dpba_get_pose_idepth_blocks(self.h, slot, n, _ptr(out)))
        return out

    def set_statuses(self, r, t, statuses):
        st = _u8(statuses)
        self._ck(self.lib.dpba_set_statuses(self.h, r, t, len(st), _ptr(st)))

    def append_statuses(self, r, t, first, statuses):
        st = _u8(statuses)
        self._ck(self.lib.dpba_append_statuses(self.h, r, t, int(first), len(st), _ptr(st)))

    def set_frame_marginalized(self, slot, is_marginalized):
        self._ck(self.lib.dpba_set_frame_marginalized(self.h, slot, int(is_marginalized)))

    def get_statuses(self, r, t):
        n = self.num_landmarks(r)
        st, cand = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.dpba_get_statuses(self.h, r, t, n, _ptr(st), _ptr(cand)))
        return st, cand

    def get_residual_scalars(self, r, t):
        """-> (energy float32[n], reprojection_jacobians_valid uint8[n]) of the residual vector r -> t."""
        n = self.num_landmarks(r)
        e, jv = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self._ck(self.lib.dpba_get_residual_scalars(self.h, r, t, n, _ptr(e), _ptr(jv)))
        return e, jv

    def set_frame_statuses(self, r, per_target):
        """per_target: dict {target slot: uint8[n]} or list indexed by slot (None entries skipped)."""
        n_fr = self.n_frames
        items = per_target.items() if isinstance(per_target, dict) else enumerate(per_target)
        arrs = {t: _u8(a) for t, a in items if a is not None and t != r}
        n = len(next(iter(arrs.values()))) if arrs else 0
        ptrs = (C.c_void_p * n_fr)(*[arrs[t].ctypes.data if t in arrs else None for t in range(n_fr)])
        self._ck(self.lib.dpba_set_frame_statuses(self.h, r, n, ptrs))

    def set_window_landmarks(self, uv, idepth, patch, flags=None, statuses=None):
        """Every slot's landmarks (lists indexed by slot) and, optionally, every residual vector's statuses
        (dict {(r, t): uint8[n_r]}) in one call: dpba_set_window_landmarks."""
        n_fr = self.n_frames
        uv, idepth, patch = [_f32(a) for a in uv], [_f32(a) for a in idepth], [_f32(a) for a in patch]
        fl = [None] * n_fr if flags is None else [_u8(a) for a in flags]
        st = {} if statuses is None else {k: _u8(v) for k, v in statuses.items()}
        cnt = np.array([len(a) for a in idepth], np.int32)
        arr = lambda xs: (C.c_void_p * len(xs))(*[None if x is None or x.size == 0 else x.ctypes.data for x in xs])  # noqa: E731
        ps = None if statuses is None else arr([st.get((r, t)) if r != t else None for r in range(n_fr) for t in range(n_fr)])
        self._ck(self.lib.dpba_set_window_landmarks(self.h, _ptr(cnt), arr(uv), arr(idepth), arr(patch),
                                                    None if flags is None else arr(fl), ps))

    def get_frame_statuses(self, r):
        """-> (statuses, candidates), each uint8 [n_frames][n]; row r is unused (zeros)."""
        n_fr, n = self.n_frames, self.num_landmarks(r)
        st, cd = np.zeros((n_fr, n), np.uint8), np.zeros((n_fr, n), np.uint8)
        b_st, b_cd = st.ctypes.data, cd.ctypes.data  # C-contiguous: row t starts n bytes after row t - 1
        ps = (C.c_void_p * n_fr)(*[b_st + t * n if t != r and n else None for t in range(n_fr)])
        pc = (C.c_void_p * n_fr)(*[b_cd + t * n if t != r and n else None for t in range(n_fr)])
        self._ck(self.lib.dpba_get_frame_statuses(self.h, r, n, ps, pc))
        return st, cd

    def set_state(self, eps=None, step=None):
        eps, step = _f64(eps), _f64(step)
        self._ck(self.lib.dpba_set_state(self.h, _ptr(eps), _ptr(step)))

    def get_state(self):
        n = BLOCK * self.n_frames
        eps, step = np.zeros(n), np.zeros(n)
        self._ck(self.lib.dpba_get_state(self.h, _ptr(eps), _ptr(step)))
        return eps, step

    def first_estimate(self):
        self._ck(self.lib.dpba_first_estimate(self.h))

    def evaluate(self, sigma, huber=True, fej=True):
        e, n = _D(), _I()
        self._ck(self.lib.dpba_evaluate(self.h, sigma, int(huber), int(fej), C.byref(e), C.byref(n)))
        return e.value, n.value

    def evaluate_jacobians(self, sigma, huber=True, fej=True):
        self._ck(self.lib.dpba_evaluate_jacobians(self.h, sigma, int(huber), int(fej)))

    def download_residual_block(self, r, t):
        n = self.num_landmarks(r)
        out = dict(
            r=np.zeros((n, 8), np.float32), J_ref=np.zeros((n, 8, 8), np.float32),
            J_tgt=np.zeros((n, 8, 8), np.float32), d_idepth=np.zeros((n, 8), np.float32),
            w=np.zeros(n, np.float32), e=np.zeros(n, np.float32), status=np.zeros(n, np.uint8),
            cand=np.zeros(n, np.uint8))
        v = ResidualView(n, _ptr(out["r"]), _ptr(out["J_ref"]), _ptr(out["J_tgt"]), _ptr(out["d_idepth"]),
                         _ptr(out["w"]), _ptr(out["e"]), _ptr(out["status"]), _ptr(out["cand"]))
        self._ck(self.lib.dpba_download_residual_block(self.h, r, t, C.byref(v)))
        return out

    def linearize(self, sigma, huber=True, fej=True, for_marginalized=False, materialized=False):
        d = BLOCK * self.n_frames
        Hp, bp, Hs, bs = np.zeros((d, d)), np.zeros(d), np.zeros((d, d)), np.zeros(d)
        fn = self.lib.dpba_linearize_materialized if materialized else self.lib.dpba_linearize
        self._ck(fn(self.h, sigma, int(huber), int(fej), int(for_marginalized), _ptr(Hp), _ptr(bp), _ptr(Hs),
                    _ptr(bs)))
        return Hp, bp, Hs, bs

    def back_substitute(self, step_pose, lam):
        s = _f64(step_pose)
        self._ck(self.lib.dpba_back_substitute(self.h, _ptr(s), lam))

    def accept(self):
        a, b = _D(), _D()
        self._ck(self.lib.dpba_accept(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reject(self):
        self._ck(self.lib.dpba_reject(self.h))

    def change_residual_statuses(self, accept=True):
        self._ck(self.lib.dpba_change_residual_statuses(self.h, int(accept)))

    def landmarks_energy(self, for_marginalized=False):
        e, n = _D(), _I()
        self._ck(self.lib.dpba_landmarks_energy(self.h, int(for_marginalized), C.byref(e), C.byref(n)))
        return e.value, n.value

    def update_point_statuses(self, min_valid, sigma):
        t = _D()
        self._ck(self.lib.dpba_update_point_statuses(self.h, min_valid, sigma, C.byref(t)))
        return t.value

    def refine_immature_landmarks(self, ref_slot, proj_xy, idepth, patch, minimum_inliers, sigma=20.0):
        """optimizeImmatureLandmark for every candidate -> (idepth, activate, n_valid)."""
        xy, idp, pt = _f32(proj_xy), _f32(idepth), _f32(patch)
        n = len(idp)
        out, act, nv = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.int32)
        self._ck(self.lib.dpba_refine_immature_landmarks(self.h, ref_slot, n, _ptr(xy), _ptr(idp), _ptr(pt), minimum_inliers,
                                                         sigma, _ptr(out), _ptr(act), _ptr(nv)))
        return out, act.astype(bool), nv

    def create_reference_depth_maps(self, n_levels=4, idepth_variance=-1.0):
        """createReferenceDepthMaps from the resident window -> list over levels of (idepth_sum, weight), (H_l, W_l) each."""
        W, H = self.cfg.width, self.cfg.height
        out = [(np.empty((H >> l, W >> l), np.float32), np.empty((H >> l, W >> l), np.float32)) for l in range(n_levels)]
        pi = (C.c_void_p * n_levels)(*[a.ctypes.data for a, _ in out])
        pw = (C.c_void_p * n_levels)(*[b.ctypes.data for _, b in out])
        self._ck(self.lib.dpba_create_reference_depth_maps(self.h, n_levels, float(idepth_variance), pi, pw))
        return out

    def solve_lm(self, sigma=20.0, ab_reg=(1e12, 1e8), fixed_reg=1e16, max_it=7, min_it=3, ftol=1e-8, ptol=1e-8,
                 force_accept=True, lambda0=1e-5, decrease=1.0, increase=1.0, fej=True, H_marg=None, b_marg=None,
                 energy_marg=0.0):
        """levenberg_marquardt_algorithm::solve entirely on the device (dpba_solve_lm)."""
        o = LmOptions(max_it, min_it, int(force_accept), int(fej), lambda0, ftol, ptol, decrease, increase, sigma,
                      (C.c_double * 2)(*ab_reg), fixed_reg)
        r = LmResult()
        Hm, bm = _f64(H_marg), _f64(b_marg)
        self._ck(self.lib.dpba_solve_lm(self.h, C.byref(o), _ptr(Hm), _ptr(bm), energy_marg, C.byref(r)))
        return r.energy, r.iterations, bool(r.converged), r.number_of_valid_residuals

    PROFILE_KINDS = ("linearize_fused", "schur", "residual_sweep", "materialise_sweep", "assemble", "back_substitute",
                     "pair_setup", "lm_step", "core_reduce", "assemble_blocks", "schur_reduce", "lm_control")

    def set_option(self, name, value):
        self._ck(self.lib.dpba_set_option(self.h, name.encode(), int(value)))

    def profile_enable(self, on=True):
        self._ck(self.lib.dpba_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms, n = np.zeros(len(self.PROFILE_KINDS)), np.zeros(len(self.PROFILE_KINDS), np.int32)
        self._ck(self.lib.dpba_profile_read(self.h, _ptr(ms), _ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    def comm_init(self, uid: bytes, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self.lib.dpba_comm_init(self.h, C.cast(buf, C.c_void_p), rank, world))

    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange mailbox (all-gather them, then peer_attach)."""
        buf = (C.c_uint8 * 64)()
        self._ck(self.lib.dpba_peer_export(self.h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def peer_barrier(self):
        """Device-side rendezvous of the attached ranks on the handle's stream (dpba_peer_barrier)."""
        self._ck(self.lib.dpba_peer_barrier(self.h))

    def peer_attach(self, handles: bytes, rank, world):
        assert len(handles) == 64 * world
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._ck(self.lib.dpba_peer_attach(self.h, C.cast(buf, C.c_void_p), rank, world))


def attach_peers(h: "Handle", rank, world, device):
    """Exchange the mailbox handles over torch.distributed and attach them (callers barrier afterwards)."""
    import torch
    import torch.distributed as dist

    mine = torch.frombuffer(bytearray(h.peer_export()), dtype=torch.uint8).to(device)
    every = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(every, mine)
    h.peer_attach(b"".join(bytes(t.cpu().numpy().tobytes()) for t in every), rank, world)
    dist.barrier()


def launch_count() -> int:
    return int(load_library().dpba_launch_count())


def comm_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_uint8 * 128)()
    rc = lib.dpba_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise DpbaError(f"dpba_comm_unique_id failed with {rc}")
    return bytes(buf)


def upload_window(win, max_frames=None, max_points=None, device=0, rank=0, world_size=1) -> Handle:
    """Create a handle and load a dsopp_b200.synth.SynthWindow into it (landmark shard `rank` of `world_size`)."""
    from .sharding import shard_indices
    n = win.n_frames
    shard = [shard_indices(len(f.idepth), rank, world_size) for f in win.frames]
    mp = max_points or max(1, max(len(s) for s in shard))
    h = Handle(max_frames or max(2, n), mp, win.width, win.height, device, rank, world_size)
    for f in win.frames:
        h.push_frame(f.frame_id, f.image, f.mask, f.T_w_lin, f.exposure, f.ab0, f.intr, f.fixed)
    for i, f in enumerate(win.frames):
        s = shard[i]
        h.set_landmarks(i, f.uv[s], f.idepth[s], f.patch[s], f.flags[s])
    for (r, t), st in win.statuses.items():
        h.set_statuses(r, t, st[shard[r]])
    h.set_state(np.concatenate([f.state_eps for f in win.frames]), np.zeros(BLOCK * n))
    return h
