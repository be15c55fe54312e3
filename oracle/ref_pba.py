"""ctypes wrapper of oracle/_ref/libdsopp_ref_pba.so -- the reference's own photometric bundle adjustment compiled from its
sources against stand-in third-party headers (oracle/build_ref_pba.py).  Test infrastructure only."""
import ctypes as C

import numpy as np

from . import build_ref_pba

P = 8
BLOCK = 8
_lib = None


def available():
    return build_ref_pba.available()


def load():
    global _lib
    if _lib is None:
        path = build_ref_pba.build()
        if path is None:
            raise RuntimeError("neither /root/reference nor a prebuilt oracle/_ref/libdsopp_ref_pba.so is present")
        lib = C.CDLL(path)
        vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
        lib.refpba_create.restype = vp
        sig = {
            "refpba_destroy": [vp],
            "refpba_add_frame": [vp, i, i64, vp, d, vp, vp, vp, i, i, vp, i, vp],
            "refpba_add_landmarks": [vp, i, i, vp, vp, vp, vp],
            "refpba_set_statuses": [vp, i, i, i, vp],
            "refpba_set_frame_state": [vp, i, vp, vp],
            "refpba_set_frame_flags": [vp, i, i, i],
            "refpba_set_idepth_steps": [vp, i, vp],
            "refpba_set_marginalized": [vp, i, vp, vp, d],
            "refpba_get_marginalized": [vp, vp, vp, vp, vp],
            "refpba_n_frames": [vp],
            "refpba_n_landmarks": [vp, i],
            "refpba_first_estimate": [vp],
            "refpba_evaluate": [vp, i, i, i, d],
            "refpba_change_statuses": [vp, i],
            "refpba_get_residuals": [vp, i, i] + [vp] * 14,
            "refpba_get_landmarks": [vp, i] + [vp] * 10,
            "refpba_get_frame_state": [vp, i, vp, vp, vp],
            "refpba_linear_systems": [vp, i, i, vp, d, vp, vp, vp, vp],
            "refpba_calculate_idepths": [vp, vp, d],
            "refpba_landmarks_energy": [vp, i, vp, vp],
            "refpba_normal_solve": [i, vp, vp, vp],
            "refpba_solve": [vp, i, i, d, d, d, i, d, vp, d, vp, vp, vp],
            "refpba_get_intensities": [vp, i, i, vp, vp],
            "refpba_marginalize": [vp, i, d, vp, d],
            "refpba_update_point_statuses": [vp, i, d],
            "refpba_relinearize_system": [vp],
            "refpba_get_relative_baseline": [vp, i, vp],
            "refpba_get_linearization_point": [vp, i, vp, vp],
        }
        for name, args in sig.items():
            getattr(lib, name).argtypes = args
        assert lib.refpba_precision_bytes() == 8
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefWindow:
    """A deque of the reference's LocalFrame objects (PBA/local_frame.hpp) and the solver state around it."""

    def __init__(self):
        self.lib = load()
        self.h = self.lib.refpba_create()
        self.n_lm = []

    def close(self):
        if self.h:
            self.lib.refpba_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- building ------------------------------------------------------------------------------------------------
    def add_frame(self, frame_id, timestamp_ns, T_w_lin, exposure, ab0, intr, intensity, mask=None, fixed=False,
                  state_eps=None):
        """`intensity` is the raw (H, W) image; the reference's PixelMap computes {I, dx, dy} itself."""
        img = _f64(intensity)
        Hh, Ww = img.shape
        T = _f64(np.asarray(T_w_lin)[:3, :4])
        ab, k = _f64(ab0), _f64(intr)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        se = None if state_eps is None else _f64(state_eps)
        idx = self.lib.refpba_add_frame(self.h, int(frame_id), int(timestamp_ns), T.ctypes.data, float(exposure),
                                        ab.ctypes.data, k.ctypes.data, img.ctypes.data, Ww, Hh,
                                        None if m is None else m.ctypes.data, int(bool(fixed)),
                                        None if se is None else se.ctypes.data)
        self.n_lm.append(0)
        return idx

    def add_landmarks(self, f, uv, idepth, patch, flags=None):
        uv, idepth, patch = _f64(uv), _f64(idepth), _f64(patch)
        n = len(idepth)
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        self.lib.refpba_add_landmarks(self.h, f, n, uv.ctypes.data, idepth.ctypes.data, patch.ctypes.data,
                                      None if fl is None else fl.ctypes.data)
        self.n_lm[f] += n

    def set_statuses(self, f, t, statuses):
        st = np.ascontiguousarray(statuses, dtype=np.uint8)
        self.lib.refpba_set_statuses(self.h, f, t, len(st), st.ctypes.data)

    def set_frame_state(self, f, state_eps=None, step=None):
        a = None if state_eps is None else _f64(state_eps)
        b = None if step is None else _f64(step)
        self.lib.refpba_set_frame_state(self.h, f, None if a is None else a.ctypes.data, None if b is None else b.ctypes.data)

    def set_frame_flags(self, f, to_marginalize=False, is_marginalized=False):
        self.lib.refpba_set_frame_flags(self.h, f, int(to_marginalize), int(is_marginalized))

    def set_idepth_steps(self, f, step):
        s = _f64(step)
        assert len(s) == self.n_lm[f]
        self.lib.refpba_set_idepth_steps(self.h, f, s.ctypes.data)

    def set_marginalized(self, H, b, energy):
        H, b = _f64(H), _f64(b)
        self.lib.refpba_set_marginalized(self.h, len(b), H.ctypes.data, b.ctypes.data, float(energy))

    @property
    def n_frames(self):
        return self.lib.refpba_n_frames(self.h)

    # ---- the reference's functions -----------------------------------------------------------------------------------
    def first_estimate(self):
        self.lib.refpba_first_estimate(self.h)

    def evaluate(self, fej, jacobians, huber=False, sigma=0.0):
        self.lib.refpba_evaluate(self.h, int(fej), int(jacobians), int(huber), float(sigma))

    def change_statuses(self, accept=True):
        self.lib.refpba_change_statuses(self.h, int(accept))

    def residuals(self, f, t):
        n = self.n_lm[f]
        out = dict(status=np.zeros(n, np.uint8), cand=np.zeros(n, np.uint8), r=np.zeros((n, P)), du_id=np.zeros((n, P)),
                   dv_id=np.zeros((n, P)), du_t=np.zeros((n, P, 6)), dv_t=np.zeros((n, P, 6)), jac_valid=np.zeros(n, np.uint8),
                   J_ref=np.zeros((n, P, BLOCK)), J_tgt=np.zeros((n, P, BLOCK)), d_idepth=np.zeros((n, P)), w=np.zeros(n),
                   e=np.zeros(n), bcs=np.zeros(n))
        order = ["status", "cand", "r", "du_id", "dv_id", "du_t", "dv_t", "jac_valid", "J_ref", "J_tgt", "d_idepth", "w", "e", "bcs"]
        self.lib.refpba_get_residuals(self.h, f, t, *[out[k].ctypes.data for k in order])
        out["jac_valid"] = out["jac_valid"].astype(bool)
        return out

    def landmarks(self, f):
        n, D = self.n_lm[f], BLOCK * self.n_frames
        out = dict(idepth=np.zeros(n), idepth_step=np.zeros(n), inv_hdd=np.zeros(n), b_d=np.zeros(n), Hpd=np.zeros((n, D)),
                   ill=np.zeros(n, np.uint8), flags=np.zeros(n, np.uint8), ref_pattern=np.zeros((n, P, 2)),
                   corrected=np.zeros((n, P)), n_inliers=np.zeros(n, np.int64))
        order = ["idepth", "idepth_step", "inv_hdd", "b_d", "Hpd", "ill", "flags", "ref_pattern", "corrected", "n_inliers"]
        self.lib.refpba_get_landmarks(self.h, f, *[out[k].ctypes.data for k in order])
        out["ill"] = out["ill"].astype(bool)
        return out

    def frame_state(self, f):
        se, st, T = np.zeros(BLOCK), np.zeros(BLOCK), np.zeros((3, 4))
        self.lib.refpba_get_frame_state(self.h, f, se.ctypes.data, st.ctypes.data, T.ctypes.data)
        return se, st, T

    def linear_systems(self, for_marginalized=False, prior=None):
        """-> (H_pose, b_pose, H_schur, b_schur); prior = (affine_reg(2), fixed_reg) adds evaluateLinearSystemPrior."""
        n = BLOCK * self.n_frames
        Hp, bp, Hs, bs = np.zeros((n, n)), np.zeros(n), np.zeros((n, n)), np.zeros(n)
        reg = _f64(prior[0]) if prior is not None else np.zeros(2)
        self.lib.refpba_linear_systems(self.h, int(for_marginalized), int(prior is not None), reg.ctypes.data,
                                       float(prior[1]) if prior is not None else 0.0, Hp.ctypes.data, bp.ctypes.data,
                                       Hs.ctypes.data, bs.ctypes.data)
        return Hp, bp, Hs, bs

    def calculate_idepths(self, step_poses, lam):
        s = _f64(step_poses)
        self.lib.refpba_calculate_idepths(self.h, s.ctypes.data, float(lam))

    def landmarks_energy(self, for_marginalized=False):
        e, n = C.c_double(), C.c_int32()
        self.lib.refpba_landmarks_energy(self.h, int(for_marginalized), C.byref(e), C.byref(n))
        return e.value, n.value

    def solve(self, fej=True, max_iterations=7, trust_region_radius=1e5, function_tolerance=1e-8, parameter_tolerance=1e-8,
              force_accept=True, sigma_huber=20.0, affine_reg=(1e12, 1e8), fixed_reg=1e12):
        """EigenPhotometricBundleAdjustment::solve up to the LM result (eigen_photometric_bundle_adjustment.cpp:56-84)."""
        reg = _f64(affine_reg)
        e, n, c = C.c_double(), C.c_int32(), C.c_int32()
        self.lib.refpba_solve(self.h, int(fej), int(max_iterations), float(trust_region_radius), float(function_tolerance),
                              float(parameter_tolerance), int(force_accept), float(sigma_huber), reg.ctypes.data,
                              float(fixed_reg), C.byref(e), C.byref(n), C.byref(c))
        return e.value, n.value, bool(c.value)

    def marginalize(self, fej=True, sigma_huber=20.0, affine_reg=(1e12, 1e8), fixed_reg=1e12):
        """The frames_.size() > 1 part of pushFrame (eigen_photometric_bundle_adjustment.cpp:121-130); frames flagged
        to_marginalize leave the deque.  -> (H_marg, b_marg, energy_marg)"""
        reg = _f64(affine_reg)
        self.lib.refpba_marginalize(self.h, int(fej), float(sigma_huber), reg.ctypes.data, float(fixed_reg))
        return self.marginalized()

    def marginalized(self):
        size, e = C.c_int(), C.c_double()
        self.lib.refpba_get_marginalized(self.h, None, None, C.byref(e), C.byref(size))
        n = size.value
        H, b = np.zeros((n, n)), np.zeros(n)
        self.lib.refpba_get_marginalized(self.h, H.ctypes.data, b.ctypes.data, C.byref(e), C.byref(size))
        return H, b, e.value

    def update_point_statuses(self, min_valid=1, sigma_huber=20.0):
        """PhotometricBundleAdjustment::updatePointStatuses (photometric_bundle_adjustment.cpp:318-406)."""
        self.lib.refpba_update_point_statuses(self.h, int(min_valid), float(sigma_huber))

    def relinearize_system(self):
        """PhotometricBundleAdjustment::relinearizeSystem (photometric_bundle_adjustment.cpp:307-316)."""
        self.lib.refpba_relinearize_system(self.h)

    def relative_baseline(self, f):
        out = np.zeros(self.n_lm[f])
        self.lib.refpba_get_relative_baseline(self.h, f, out.ctypes.data)
        return out

    def linearization_point(self, f):
        T, ab = np.zeros((3, 4)), np.zeros(2)
        self.lib.refpba_get_linearization_point(self.h, f, T.ctypes.data, ab.ctypes.data)
        return T, ab

    def get_intensities(self, f, uv):
        uv = _f64(uv)
        n = len(uv)
        out = np.zeros((n, P))
        self.lib.refpba_get_intensities(self.h, f, n, uv.ctypes.data, out.ctypes.data)
        return out


def normal_solve(H, b):
    """NormalLinearSystem<>::solve (normal_linear_system.cpp:52-60)."""
    lib = load()
    H, b = _f64(H), _f64(b)
    x = np.zeros(len(b))
    lib.refpba_normal_solve(len(b), H.ctypes.data, b.ctypes.data, x.ctypes.data)
    return x


def window_from_synth(win, raw_images):
    """Fill a RefWindow from a dsopp_b200.synth.SynthWindow; raw_images[k] is the (H, W) float64 intensity of frame k."""
    rw = RefWindow()
    for k, f in enumerate(win.frames):
        rw.add_frame(f.frame_id, f.timestamp, f.T_w_lin, f.exposure, f.ab0, f.intr, raw_images[k], f.mask, f.fixed, f.state_eps)
        rw.add_landmarks(k, f.uv, f.idepth, f.patch, f.flags)
    for (r, t), st in win.statuses.items():
        rw.set_statuses(r, t, st)
    return rw


def pose_alignment_solve(ref_T, ref_exposure, ref_ab, ref_intensity, tgt_T, tgt_exposure, tgt_ab, tgt_intensity, intr,
                         idepth_sum, weight, tgt_mask=None, sigma_huber=20.0, affine_reg=(1e12, 1e8), max_iterations=50,
                         trust_region_radius=1e2, function_tolerance=1e-5, parameter_tolerance=1e-5, prior_rotation=None):
    """The reference's coarse-tracker alignment: the depth-map LocalFrame constructor (local_frame.hpp:350-393), class
    PoseAlignerProblem (eigen_pose_alignment.cpp:24-242) under levenberg_marquardt_algorithm::solve, set up as
    EigenPoseAlignment::solve does (:275-329).  Images are raw (H, W) intensities; (idepth_sum, weight) is the depth map.
    -> dict(energy, n_valid, converged, T_t_r (4x4), ab_eps, H (8x8), uv, idepth, patch of the landmarks it made)"""
    lib = load()
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.refpa_solve.restype = d
    lib.refpa_solve.argtypes = [vp, d, vp, vp, vp, d, vp, vp, vp, vp, i, i, vp, vp, d, vp, i, d, d, d, vp, vp, vp, vp, vp, vp, vp, i,
                                vp, vp, vp]
    ri, ti = _f64(ref_intensity), _f64(tgt_intensity)
    Hh, Ww = ri.shape
    rT, tT = _f64(np.asarray(ref_T)[:3, :4]), _f64(np.asarray(tgt_T)[:3, :4])
    rab, tab, k, reg = _f64(ref_ab), _f64(tgt_ab), _f64(intr), _f64(affine_reg)
    ds, wt = _f64(idepth_sum), _f64(weight)
    m = None if tgt_mask is None else np.ascontiguousarray(tgt_mask, dtype=np.uint8)
    pr = None if prior_rotation is None else _f64(prior_rotation)
    cap = Hh * Ww
    T, ab, H = np.zeros((3, 4)), np.zeros(2), np.zeros((8, 8))
    uv, idp, pt = np.zeros((cap, 2)), np.zeros(cap), np.zeros(cap)
    nv, cv, nl = C.c_int32(), C.c_int32(), C.c_int32()
    e = lib.refpa_solve(rT.ctypes.data, float(ref_exposure), rab.ctypes.data, ri.ctypes.data, tT.ctypes.data,
                        float(tgt_exposure), tab.ctypes.data, ti.ctypes.data, None if m is None else m.ctypes.data,
                        k.ctypes.data, Ww, Hh, ds.ctypes.data, wt.ctypes.data, float(sigma_huber), reg.ctypes.data,
                        int(max_iterations), float(trust_region_radius), float(function_tolerance), float(parameter_tolerance),
                        None if pr is None else pr.ctypes.data, T.ctypes.data, ab.ctypes.data, H.ctypes.data,
                        C.addressof(nv), C.addressof(cv), C.addressof(nl), cap, uv.ctypes.data, idp.ctypes.data, pt.ctypes.data)
    n = nl.value
    return dict(energy=e, n_valid=nv.value, converged=bool(cv.value), T_t_r=np.vstack([T, [0, 0, 0, 1.0]]), ab_eps=ab, H=H,
                uv=uv[:n].copy(), idepth=idp[:n].copy(), patch=pt[:n].copy())


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def photometric_correction(gray_u8, lut, vignetting_u8=None):
    """The reference's photometricallyCorrectedImage (photometrically_corrected_image.cpp:9-29) -> (H, W) float64."""
    lib = load()
    g, v, t = _u8(gray_u8), _u8(vignetting_u8), _f64(lut)
    H, W = g.shape
    out = np.zeros((H, W))
    lib.refpyr_photometric_correction.restype = None
    lib.refpyr_photometric_correction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.refpyr_photometric_correction(g.ctypes.data, H, W, t.ctypes.data, None if v is None else v.ctypes.data, out.ctypes.data)
    return out


def downscale(image):
    """The reference's downscaleImage (downscale_image.hpp:16-33) -> (H / 2, W / 2) float64."""
    lib = load()
    im = _f64(image)
    H, W = im.shape
    out = np.zeros((H // 2, W // 2))
    lib.refpyr_downscale.restype = None
    lib.refpyr_downscale.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.refpyr_downscale(im.ctypes.data, H, W, out.ctypes.data)
    return out


def pixel_data_frame(gray_u8, lut, vignetting_u8, levels):
    """The reference's PixelDataFrame constructor (pixel_data_frame.cpp:12-31): photometric correction, levels - 1 halvings,
    every level packed by PixelMap<1> -> list of (H >> l, W >> l, 3) float64 arrays {I, dx, dy}."""
    lib = load()
    g, v, t = _u8(gray_u8), _u8(vignetting_u8), _f64(lut)
    H, W = g.shape
    sizes = [(H >> l, W >> l) for l in range(min(levels, 5))]
    out = np.zeros(sum(h * w * 3 for h, w in sizes))
    lib.refpyr_pixel_data_frame.restype = C.c_int
    lib.refpyr_pixel_data_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    n = lib.refpyr_pixel_data_frame(g.ctypes.data, H, W, t.ctypes.data, None if v is None else v.ctypes.data, int(levels),
                                    out.ctypes.data)
    assert n == len(sizes), (n, sizes)
    res, k = [], 0
    for h, w in sizes:
        res.append(out[k:k + h * w * 3].reshape(h, w, 3).copy())
        k += h * w * 3
    return res
