"""ctypes binding of the C++ host side (libdsopp_pba_host.so): CudaPhotometricBundleAdjustment, the LM driver and
NormalLinearSystem.  Test / bench plumbing -- a C++ caller links the headers in dsopp_b200/csrc/host/ directly."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import BLOCK, DpbaError, _f32, _f64, _ptr, _u8, pose34

_P, _I, _D, _LL = C.c_void_p, C.c_int, C.c_double, C.c_longlong
_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    capi.load_library()  # RTLD_GLOBAL: the host library resolves dpba_* from it
    import os
    if not os.path.exists(capi.HOST_LIB_PATH):
        raise DpbaError(f"{capi.HOST_LIB_PATH} is not built: run `python -m dsopp_b200.build`")
    lib = C.CDLL(capi.HOST_LIB_PATH)
    lib.dpbah_last_error.restype = C.c_char_p
    lib.dpbah_create.restype = _P
    lib.dpbah_create.argtypes = [_I] * 9 + [_D] * 7 + [_I]
    lib.dpbah_destroy.argtypes = [_P]
    lib.dpbah_handle.restype = _P
    lib.dpbah_handle.argtypes = [_P]
    lib.dpbah_push_frame.argtypes = [_P, _I, _LL, _P, _D, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]
    lib.dpbah_update_local_frame.argtypes = [_P, _I, _LL, _P, _D, _P, _P, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P]
    lib.dpbah_solve.argtypes = [_P, C.POINTER(_D), C.POINTER(_I)]
    lib.dpbah_num_frames.argtypes = [_P]
    lib.dpbah_frame_ids.argtypes = [_P, _P]
    lib.dpbah_update_frame.argtypes = [_P, _LL, _P, _P, _I, _P, _P, _P, _P, _P]
    lib.dpbah_marginalized_system.argtypes = [_P, _P, _P, C.POINTER(_D)]
    lib.dpbah_covariance.argtypes = [_P, _I, _I, _P]
    lib.dpbah_lm_solve.argtypes = [_P, _I, _P, _P, _D, _D, _D, _D, _I, _I, _D, _D, _I, _D, _D, _D, C.POINTER(_D),
                                   C.POINTER(_I)]
    lib.dpbah_normal_solve.argtypes = [_I, _P, _P, _P]
    lib.dpbah_lm_scripted.restype = _I
    lib.dpbah_lm_scripted.argtypes = [_I, _D, _D, _D, _I, _I, _D, _D, _P, _P, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P]
    lib.dpbah_reduce_system.argtypes = [_I, _P, _P, _I, _P]
    lib.dpbah_sym_pinv.argtypes = [_I, _P, _I, _P]
    _lib = lib
    return lib


def _ck(rc):
    if rc < 0:
        raise DpbaError(f"host error {rc}: {load_library().dpbah_last_error().decode()}")
    return rc


def normal_solve(H, b):
    lib = load_library()
    H, b = _f64(H), _f64(b)
    x = np.zeros_like(b)
    lib.dpbah_normal_solve(len(b), _ptr(H), _ptr(b), _ptr(x))
    return x


LM_CALL_NAMES = ["energy", "linearize", "step", "accept", "reject"]


def lm_scripted(energies, valid, norms, max_it=50, lambda0=1e-5, ftol=1e-8, ptol=1e-8, force_accept=False, min_it=0,
                dec=2.0, inc=10.0):
    """The C++ host LM driver on a scripted problem -> (calls, lambdas, energy, n_valid, converged); no GPU involved."""
    lib = load_library()
    e = _f64(energies)
    v = np.ascontiguousarray(valid, dtype=np.int32)
    nr = _f64(np.asarray(norms, dtype=np.float64).reshape(-1, 2))
    calls = np.zeros(8 * (max_it + 2), np.int32)
    lams = np.zeros(max_it + 2)
    out_e, out_v, out_c = np.zeros(1), np.zeros(1, np.int32), np.zeros(1, np.int32)
    n = lib.dpbah_lm_scripted(int(max_it), lambda0, ftol, ptol, int(force_accept), int(min_it), dec, inc, _ptr(e), _ptr(v),
                              len(e), _ptr(nr), len(nr), _ptr(calls), len(calls), _ptr(lams), len(lams), _ptr(out_e),
                              _ptr(out_v), _ptr(out_c))
    calls = calls[:n]
    return ([LM_CALL_NAMES[c] for c in calls], lams[:int((calls == 2).sum())].copy(), float(out_e[0]), int(out_v[0]),
            bool(out_c[0]))


def reduce_system(H, b, elim):
    lib = load_library()
    H, b = _f64(H).copy(), _f64(b).copy()
    e = np.ascontiguousarray(elim, dtype=np.int32)
    n = lib.dpbah_reduce_system(len(b), _ptr(H), _ptr(b), len(e), _ptr(e))
    return H.reshape(-1)[: n * n].reshape(n, n).copy(), b[:n].copy()


def sym_pinv(A, n_null):
    lib = load_library()
    A = _f64(A)
    out = np.zeros_like(A)
    lib.dpbah_sym_pinv(A.shape[0], _ptr(A), n_null, _ptr(out))
    return out


def lm_solve(handle: capi.Handle, ab0, fixed, sigma=20.0, ab_reg=(1e12, 1e8), fixed_reg=1e16, max_it=7, min_it=3,
             ftol=1e-8, ptol=1e-8, force_accept=True, lambda0=1e-5, decrease=1.0, increase=1.0):
    """levenberg_marquardt_algorithm::solve (C++) over a window already uploaded through the C ABI."""
    lib = load_library()
    ab0 = _f64(np.asarray(ab0).reshape(-1))
    fx = np.ascontiguousarray(fixed, dtype=np.int32)
    e, it = _D(), _I()
    _ck(lib.dpbah_lm_solve(handle.h, len(fx), _ptr(ab0), _ptr(fx), sigma, ab_reg[0], ab_reg[1], fixed_reg, max_it,
                           min_it, ftol, ptol, int(force_accept), lambda0, decrease, increase, C.byref(e), C.byref(it)))
    return e.value, it.value


class _WindowIOStruct(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("frame_ids", _P), ("images", _P), ("masks", _P), ("T_w_lin", _P), ("exposure", _P),
                ("ab0", _P), ("intr", _P), ("fixed", _P), ("n_landmarks", _P), ("uv", _P), ("idepth", _P), ("patch", _P),
                ("flags", _P), ("statuses", _P), ("eps0", _P), ("lm", capi.LmOptions), ("eps_out", _P), ("idepth_out", _P),
                ("inv_hdd_out", _P), ("rel_baseline_out", _P), ("flags_out", _P), ("n_inliers_out", _P),
                ("statuses_out", _P), ("energy", C.c_double), ("iterations", C.c_int32), ("n_valid", C.c_int32),
                ("converged", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("raw_gray", _P),
                ("photometric_lut", _P), ("phase_ms", C.c_double * 5), ("sync_phases", C.c_int32), ("image_channels", C.c_int32)]


class WindowStep:
    """dpbah_solve_window: one whole solver step from host buffers in ONE C++ call (push every keyframe, landmarks,
    statuses, state -> firstEstimateJacobians + device LM -> results back).  `frames`: list of dicts with the host arrays
    (frame_id, image HxWx3 f32 -- or HxW f32 intensity planes, what PixelMap::data() holds: the device then builds {I,dx,dy}
    itself --, mask HxW u8 or None, T_w_lin, exposure, ab0, intr, fixed, uv, idepth, patch, flags);
    `statuses[(r, t)]`: uint8 per landmark of r.  Arrays are used in place -- pass page-locked ones to have them DMA'd
    without a staging copy.  Results land in `self.out` (eps, and per frame idepth / inv_hdd / rel_baseline / flags /
    n_inliers / statuses[t])."""

    def __init__(self, handle: capi.Handle, frames, statuses, eps0, sigma=20.0, ab_reg=(1e12, 1e8), fixed_reg=1e16, max_it=7,
                 min_it=3, ftol=1e-8, ptol=1e-8, force_accept=True, lambda0=1e-5, decrease=1.0, increase=1.0, fej=True,
                 alloc=None, raw_gray=None, photometric_lut=None):
        self.lib = load_library()
        self.lib.dpbah_solve_window.argtypes = [_P, C.POINTER(_WindowIOStruct), _I, _I]
        self.h = handle
        n = len(frames)
        alloc = alloc or (lambda shape, dtype: np.zeros(shape, dtype))
        k = self._keep = {}
        k["ids"] = np.ascontiguousarray([f["frame_id"] for f in frames], dtype=np.int32)
        k["T"] = np.ascontiguousarray(np.stack([pose34(f["T_w_lin"]) for f in frames]))
        k["exp"] = np.ascontiguousarray([f["exposure"] for f in frames], dtype=np.float64)
        k["ab0"] = np.ascontiguousarray(np.stack([_f64(f["ab0"]) for f in frames]))
        k["intr"] = np.ascontiguousarray(np.stack([_f64(f["intr"]) for f in frames]))
        k["fixed"] = np.ascontiguousarray([int(f["fixed"]) for f in frames], dtype=np.int32)
        k["n_lm"] = np.ascontiguousarray([len(f["idepth"]) for f in frames], dtype=np.int32)
        k["eps0"] = _f64(eps0)
        k["frames"] = frames
        k["statuses"] = statuses

        def ptr_array(items):
            return (C.c_void_p * len(items))(*[None if a is None else a.ctypes.data for a in items])

        for name, dt in (("image", np.float32), ("uv", np.float32), ("idepth", np.float32), ("patch", np.float32),
                         ("flags", np.uint8)):
            for f in frames:
                a = f[name]
                assert a.dtype == dt and a.flags["C_CONTIGUOUS"], name
        k["p_images"] = ptr_array([f["image"] for f in frames])
        planes = [f["image"].ndim == 2 for f in frames]
        assert all(planes) or not any(planes), "all frames as {I,dx,dy} records or all as intensity planes"
        k["p_masks"] = ptr_array([f.get("mask") for f in frames])
        k["p_uv"] = ptr_array([f["uv"] for f in frames])
        k["p_idepth"] = ptr_array([f["idepth"] for f in frames])
        k["p_patch"] = ptr_array([f["patch"] for f in frames])
        k["p_flags"] = ptr_array([f["flags"] for f in frames])
        k["p_status"] = ptr_array([None if r == t else statuses[(r, t)] for r in range(n) for t in range(n)])
        self.out = dict(eps=np.zeros(8 * n))
        for name, dt in (("idepth", np.float32), ("inv_hdd", np.float32), ("rel_baseline", np.float32), ("flags", np.uint8),
                         ("n_inliers", np.uint32)):
            self.out[name] = [alloc(int(m), dt) for m in k["n_lm"]]
            k["p_out_" + name] = ptr_array(self.out[name])
        self.out["statuses"] = [[None if r == t else alloc(int(k["n_lm"][r]), np.uint8) for t in range(n)] for r in range(n)]
        k["p_out_status"] = ptr_array([self.out["statuses"][r][t] for r in range(n) for t in range(n)])
        io = self.io = _WindowIOStruct()
        io.n_frames = n
        io.frame_ids, io.T_w_lin, io.exposure = k["ids"].ctypes.data, k["T"].ctypes.data, k["exp"].ctypes.data
        io.ab0, io.intr, io.fixed, io.n_landmarks = (k["ab0"].ctypes.data, k["intr"].ctypes.data, k["fixed"].ctypes.data,
                                                     k["n_lm"].ctypes.data)
        cast = lambda a: C.cast(a, C.c_void_p)
        io.images, io.masks, io.uv, io.idepth = cast(k["p_images"]), cast(k["p_masks"]), cast(k["p_uv"]), cast(k["p_idepth"])
        io.patch, io.flags, io.statuses, io.eps0 = cast(k["p_patch"]), cast(k["p_flags"]), cast(k["p_status"]), k["eps0"].ctypes.data
        io.lm = capi.LmOptions(max_it, min_it, int(force_accept), int(fej), lambda0, ftol, ptol, decrease, increase, sigma,
                               (C.c_double * 2)(*ab_reg), fixed_reg)
        io.eps_out = self.out["eps"].ctypes.data
        io.idepth_out, io.inv_hdd_out = cast(k["p_out_idepth"]), cast(k["p_out_inv_hdd"])
        io.rel_baseline_out, io.flags_out = cast(k["p_out_rel_baseline"]), cast(k["p_out_flags"])
        io.n_inliers_out, io.statuses_out = cast(k["p_out_n_inliers"]), cast(k["p_out_status"])
        io.image_channels = 1 if all(planes) else 3
        if raw_gray is not None:  # 8-bit frames: dpba_push_frame_raw (photometric table + {I,dx,dy} on the device)
            k["raw"], k["lut"] = [_u8(g) for g in raw_gray], _f32(photometric_lut)
            k["p_raw"] = ptr_array(k["raw"])
            io.raw_gray, io.photometric_lut = cast(k["p_raw"]), k["lut"].ctypes.data

    def run(self):
        """-> (energy, iterations); results in self.out, byte counters in self.io.h2d_bytes / d2h_bytes."""
        _ck(self.lib.dpbah_solve_window(self.h.h, C.byref(self.io), self.h.cfg.width, self.h.cfg.height))
        return self.io.energy, self.io.iterations

    def run_sliding(self):
        """dpbah_solve_sliding: the oldest keyframe leaves, ONE keyframe arrives from the host buffers (the one that left,
        so the window keeps its n frames in rotating order), solve, results of every frame back.  Call run() once first."""
        if not hasattr(self, "_order"):
            self._order = np.arange(self.io.n_frames, dtype=np.int32)
            self.lib.dpbah_solve_sliding.argtypes = [_P, C.POINTER(_WindowIOStruct), _P, _I, _I]
        _ck(self.lib.dpbah_solve_sliding(self.h.h, C.byref(self.io), self._order.ctypes.data, self.h.cfg.width,
                                         self.h.cfg.height))
        return self.io.energy, self.io.iterations


class CudaPhotometricBundleAdjustment:
    """Python proxy of the C++ class of the same name (csrc/host/cuda_photometric_bundle_adjustment.hpp)."""

    def __init__(self, width, height, max_frames=9, max_points=4096, device=0, estimate_uncertainty=True,
                 force_accept=True, max_iterations=7, min_iterations=3, radius=1e5, ftol=1e-8, ptol=1e-8,
                 ab_reg=(1e12, 1e8), fixed_reg=1e16, sigma=20.0, device_lm=True):
        self.lib = load_library()
        self.s = self.lib.dpbah_create(width, height, max_frames, max_points, device, int(estimate_uncertainty),
                                       int(force_accept), max_iterations, min_iterations, radius, ftol, ptol,
                                       ab_reg[0], ab_reg[1], fixed_reg, sigma, int(device_lm))
        if not self.s:
            raise DpbaError("dpbah_create: " + self.lib.dpbah_last_error().decode())
        self.s = C.c_void_p(self.s)
        self._keep = []

    def close(self):
        if getattr(self, "s", None):
            self.lib.dpbah_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _statuses(n, other_ids, ref_statuses, tgt_statuses):
        ids = np.ascontiguousarray(other_ids, dtype=np.int32)
        ref = None
        if ref_statuses is not None:
            ref = np.ascontiguousarray(np.stack([_u8(ref_statuses[i]) for i in other_ids]) if len(ids) else
                                       np.zeros((0, n), np.uint8))
        tgt, cnt = None, None
        if tgt_statuses is not None:
            parts = [_u8(tgt_statuses[i]) for i in other_ids]
            cnt = np.ascontiguousarray([len(p) for p in parts], dtype=np.int32)
            tgt = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0, np.uint8))
        return ids, ref, tgt, cnt

    def push_frame(self, frame_id, timestamp, T_w_agent, exposure, ab, intr, image, mask, uv, idepth, patch, flags=None,
                   fixed=False, is_marginalized=False, other_ids=(), ref_statuses=None, tgt_statuses=None):
        T, ab, intr = pose34(T_w_agent), _f64(ab), _f64(intr)
        image, mask = _f32(image), _u8(mask)
        uv, idepth, patch, flags = _f32(uv), _f32(idepth), _f32(patch), _u8(flags)
        ids, ref, tgt, cnt = self._statuses(len(idepth), list(other_ids), ref_statuses, tgt_statuses)
        _ck(self.lib.dpbah_push_frame(self.s, frame_id, timestamp, _ptr(T), exposure, _ptr(ab), _ptr(intr), _ptr(image),
                                      _ptr(mask), int(is_marginalized), len(idepth), _ptr(uv), _ptr(idepth), _ptr(patch),
                                      _ptr(flags), int(fixed), len(ids), _ptr(ids), _ptr(ref), _ptr(tgt), _ptr(cnt)))

    def update_local_frame(self, frame_id, timestamp, T_w_agent, exposure, ab, intr, uv, idepth, patch, flags=None,
                           is_marginalized=False, other_ids=(), ref_statuses=None, tgt_statuses=None):
        T, ab, intr = pose34(T_w_agent), _f64(ab), _f64(intr)
        uv, idepth, patch, flags = _f32(uv), _f32(idepth), _f32(patch), _u8(flags)
        ids, ref, tgt, cnt = self._statuses(len(idepth), list(other_ids), ref_statuses, tgt_statuses)
        _ck(self.lib.dpbah_update_local_frame(self.s, frame_id, timestamp, _ptr(T), exposure, _ptr(ab), _ptr(intr),
                                              int(is_marginalized), len(idepth), _ptr(uv), _ptr(idepth), _ptr(patch),
                                              _ptr(flags), len(ids), _ptr(ids), _ptr(ref), _ptr(tgt), _ptr(cnt)))

    def solve(self):
        e, it = _D(), _I()
        _ck(self.lib.dpbah_solve(self.s, C.byref(e), C.byref(it)))
        return e.value, it.value

    def marginalize_now(self):
        """updateMarginalizedLinearSystem of the flagged landmarks / frames (what the next pushFrame does first)."""
        self.lib.dpbah_marginalize_now.argtypes = [_P]
        _ck(self.lib.dpbah_marginalize_now(self.s))

    @property
    def frame_ids(self):
        ids = np.zeros(capi.MAX_FRAMES, np.int32)
        n = self.lib.dpbah_frame_ids(self.s, _ptr(ids))
        return ids[:n].tolist()

    def update_frame(self, timestamp, n):
        T, ab = np.zeros(12), np.zeros(2)
        out = dict(idepth=np.zeros(n, np.float32), variance=np.zeros(n, np.float32), baseline=np.zeros(n, np.float32),
                   outlier=np.zeros(n, np.uint8), inliers=np.zeros(n, np.uint32))
        _ck(self.lib.dpbah_update_frame(self.s, timestamp, _ptr(T), _ptr(ab), n, _ptr(out["idepth"]),
                                        _ptr(out["variance"]), _ptr(out["baseline"]), _ptr(out["outlier"]),
                                        _ptr(out["inliers"])))
        out["T_w_agent"], out["ab"] = T.reshape(3, 4), ab
        return out

    def marginalized_system(self):
        n = BLOCK * self.lib.dpbah_num_frames(self.s)
        H, b, e = np.zeros((n, n)), np.zeros(n), _D()
        self.lib.dpbah_marginalized_system(self.s, _ptr(H), _ptr(b), C.byref(e))
        return H, b, e.value

    def covariance(self, ref_id, tgt_id):
        out = np.zeros(36)
        if self.lib.dpbah_covariance(self.s, ref_id, tgt_id, _ptr(out)) != 0:
            return None
        return out.reshape(6, 6)


class PoseAligner:
    """C++ CudaPoseAlignment (csrc/host/cuda_pose_alignment.hpp) through libdsopp_pba_host.so."""

    def __init__(self, max_width, max_height, max_iterations=50, sigma=20.0, ab_reg=(1e12, 1e8)):
        self.lib = load_library()
        self.lib.dpah_create.restype = _P
        self.lib.dpah_create.argtypes = [_I, _I, _I, _D, _D, _D]
        self.lib.dpah_destroy.argtypes = [_P]
        self.lib.dpah_align_level.restype = _D
        self.lib.dpah_align_level.argtypes = [_P, _I, _I, _P, _P, _P, _P, _P, _D, _P, _P, _P, _P, _D, _P, _P, _P, _P, _P, _P]
        self.s = self.lib.dpah_create(max_width, max_height, max_iterations, sigma, ab_reg[0], ab_reg[1])
        if not self.s:
            raise DpbaError(f"dpah_create failed: {self.lib.dpah_last_error().decode() if hasattr(self.lib, 'dpah_last_error') else self.lib.dpbah_last_error().decode()}")

    def close(self):
        if self.s:
            self.lib.dpah_destroy(self.s)
            self.s = None

    def align_level(self, intr, ref_image, idepth_sum, weight, ref_T, ref_exposure, ref_ab, tgt_image, tgt_mask, tgt_T_guess,
                    tgt_exposure, tgt_ab, prior_rotation=None):
        """reset(); pushFrame(reference + depth map); pushFrame(target); solve() -> dict(rmse, T_w_target, ab, cov, n)."""
        ref_image, tgt_image = _f32(ref_image), _f32(tgt_image)
        ids, w, mask = _f32(idepth_sum), _f32(weight), _u8(tgt_mask)
        H, W = w.shape
        it, rT, tT = _f64(intr), pose34(ref_T), pose34(tgt_T_guess)
        rab, tab, pr = _f64(ref_ab), _f64(tgt_ab), _f64(prior_rotation)
        T, ab, cov, n = np.zeros(12), np.zeros(2), np.zeros(36), _I()
        rmse = self.lib.dpah_align_level(self.s, W, H, _ptr(it), _ptr(ref_image), _ptr(ids), _ptr(w), _ptr(rT), ref_exposure,
                                         _ptr(rab), _ptr(tgt_image), _ptr(mask), _ptr(tT), tgt_exposure, _ptr(tab), _ptr(pr),
                                         _ptr(T), _ptr(ab), _ptr(cov), C.byref(n))
        if rmse == -2.0:
            raise DpbaError(self.lib.dpbah_last_error().decode())
        return dict(rmse=rmse, T_w_target=np.vstack([T.reshape(3, 4), [0, 0, 0, 1.0]]), ab=ab, cov=cov.reshape(6, 6), n=n.value)
