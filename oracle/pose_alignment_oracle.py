"""CPU oracle (NumPy, float64) of DSOPP's coarse-tracker direct image alignment.

TEST INFRASTRUCTURE ONLY (see oracle/pba_oracle.py).  PARITY PINNED against the reference itself: its depth-map LocalFrame
constructor and its class PoseAlignerProblem are compiled from their sources (oracle/build_ref_pba.py,
oracle/ref_shims/ref_pose_alignment.cpp) and run under the reference's LM driver; landmark list exactly, energy / pose /
affine increment / Hessian at 1e-9 (tests/test_reference_pba.py, golden vectors in tests/golden/ref_pba.npz).  Also held by
property tests (tests/test_pose_alignment_oracle.py: finite-difference Jacobians, recovery of a known relative pose as
test_ceres_pose_alignment.cpp:100-139 does).

Restates, paths relative to /root/reference/src/:
  PoseAlignerProblem                energy/problems/src/eigen_pose_alignment.cpp:28-241
  EigenPoseAlignment::solve         energy/problems/src/eigen_pose_alignment.cpp:275-329
  depth-map LocalFrame constructor  energy/problems/internal/energy/problems/photometric_bundle_adjustment/local_frame.hpp:350-393
  factory constants                 tracker/tracker/src/fabric.cpp:123-147
The instantiation is the tracker's: PoseAlignment<SE3, Pinhole, PatternSize = 1, PixelMap, C = 1>,
OPTIMIZE_AFFINE_BRIGHTNESS = true (tracker/tracker/include/tracker/monocular/monocular_tracker.hpp:128-130,
energy/problems/include/energy/problems/pose_alignment/eigen_pose_alignment.hpp:25).
"""
from __future__ import annotations

import numpy as np

from . import pba_oracle as O

NUM_PARAMETERS = 8  # Motion::DoF + 2, eigen_pose_alignment.cpp:31


class PAFrame:
    """The fields of LocalFrame the aligner reads (one pyramid level)."""

    def __init__(self, T_w_agent, exposure, ab0, intr, image, mask, timestamp=0):
        self.T_lin = np.array(T_w_agent, dtype=np.float64)
        self.exposure = float(exposure)
        self.ab0 = np.array(ab0, dtype=np.float64)
        self.intr = np.array(intr, dtype=np.float64)
        self.image = np.asarray(image, dtype=np.float64)  # (H, W, 3) {I, dx, dy}
        self.H, self.W = self.image.shape[:2]
        self.mask = np.asarray(mask)
        self.timestamp = timestamp


def landmarks_from_depth_map(idepth_sum, weight, image):
    """LocalFrame depth-map constructor, local_frame.hpp:367-392: every pixel inside the 4-px border with weight > 0
    and idepth / weight >= 1e-6 becomes a 1-pixel landmark {(x, y), idepth, patch = I(x, y)}; y outer, x inner."""
    Hh, Ww = weight.shape
    k = 4
    ys, xs = np.nonzero(weight[k:Hh - k, k:Ww - k] > 0)
    ys, xs = ys + k, xs + k
    idepth = idepth_sum[ys, xs] / weight[ys, xs]
    keep = idepth >= 1e-6
    ys, xs, idepth = ys[keep], xs[keep], idepth[keep]
    uv = np.stack([xs, ys], axis=1).astype(np.float64)
    patch = np.asarray(image, dtype=np.float64)[ys, xs, 0]  # PatternPatch::getIntensities at an integer pixel
    return uv, idepth, patch


def mean_square_optical_flow(idepth_sum, weight, T_t_r, intr):
    """calculateMeanSquareOpticalFlow, src/tracker/tracker/src/monocular_tracker.cpp:104-133 (the keyframe decision's
    input, evaluated on level 0 of the reference depth map with the pose the aligner returned, :474-480): over the same
    pixels the depth-map LocalFrame takes (4-px border, weight > 0, idepth >= 1e-6) that reproject successfully,
    sqrt(mean |unproject(x) - unproject(reprojection)|^2).  Empty set -> NaN (0 / 0), as in the reference."""
    uv, idepth, _ = landmarks_from_depth_map(idepth_sum, weight, np.zeros(weight.shape + (1,)))
    Hh, Ww = weight.shape
    fx, fy, cx, cy = intr
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    Kinv = np.eye(4)
    Kinv[0, 0], Kinv[1, 1], Kinv[0, 2], Kinv[1, 2] = 1 / fx, 1 / fy, -cx / fx, -cy / fy
    A = K @ np.asarray(T_t_r, dtype=np.float64)[:3, :4] @ Kinv   # ArrayReprojector::reproject_, camera_reproject.hpp:256
    p = uv @ A[:, :2].T + A[:, 2] + idepth[:, None] * A[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = p[:, :2] / p[:, 2:3]

    def roi(q):
        return (q[:, 0] >= 4) & (q[:, 1] >= 4) & (q[:, 0] <= Ww - 5) & (q[:, 1] <= Hh - 5)

    with np.errstate(invalid="ignore"):
        ok = (idepth > -1e-4) & (idepth < 1010.0) & roi(uv) & (p[:, 2] > 0) & roi(np.where(np.isfinite(t), t, -1.0))
    d = (uv[ok] - t[ok]) / np.array([fx, fy])    # rays have z = 1 (pinhole_camera.hpp:137-139): the z difference is 0
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.sqrt(np.float64((d * d).sum()) / np.float64(ok.sum()))), int(ok.sum())


def mask_valid_checked(mask, pts):
    """CameraMask::valid<true> (bounds checked), sensors/.../mask/camera_mask.hpp:48-66."""
    xi = np.floor(np.abs(pts[..., 0]) + 0.5).astype(np.int64) * np.sign(pts[..., 0]).astype(np.int64)
    yi = np.floor(np.abs(pts[..., 1]) + 0.5).astype(np.int64) * np.sign(pts[..., 1]).astype(np.int64)
    inb = (xi >= 0) & (xi < mask.shape[1]) & (yi >= 0) & (yi < mask.shape[0])
    out = np.zeros(pts.shape[:-1], dtype=bool)
    out[inb] = mask[yi[inb], xi[inb]] != 0
    return out.all(axis=-1)


class PoseAlignerProblem:
    """eigen_pose_alignment.cpp:28-241."""

    def __init__(self, ref: PAFrame, tgt: PAFrame, uv, idepth, patch, sigma_huber, ab_reg, T_t_r, ab_eps=None):
        self.ref, self.tgt = ref, tgt
        self.uv = np.asarray(uv, dtype=np.float64).reshape(-1, 1, 2)  # reference_pattern of a 1-pixel pattern
        self.idepth = np.asarray(idepth, dtype=np.float64)
        self.patch = np.asarray(patch, dtype=np.float64)
        self.sigma = float(sigma_huber)
        self.ab_reg = np.asarray(ab_reg, dtype=np.float64)
        self.T = np.array(T_t_r, dtype=np.float64)
        self.ab_eps = np.zeros(2) if ab_eps is None else np.array(ab_eps, dtype=np.float64)
        self.old_T, self.old_ab_eps = self.T.copy(), self.ab_eps.copy()
        self.H = np.zeros((NUM_PARAMETERS, NUM_PARAMETERS))
        self.b = np.zeros(NUM_PARAMETERS)
        self.step = np.zeros(NUM_PARAMETERS)
        n = len(self.idepth)
        self.success = np.zeros(n, dtype=bool)
        self.t_patch = np.zeros(n)
        self.dI_u = np.zeros(n)
        self.dI_v = np.zeros(n)

    def _scale(self):
        ab_t = self.tgt.ab0 + self.ab_eps
        return (self.tgt.exposure / self.ref.exposure) * np.exp(ab_t[0] - self.ref.ab0[0]), ab_t

    def calculate_energy(self):  # :55-108
        s, ab_t = self._scale()
        rp = O.Reprojector(self.ref, self.tgt, self.T)
        tp, ok = rp.values(self.uv, self.idepth)
        ok = ok & mask_valid_checked(self.tgt.mask, np.where(ok[:, None, None], tp, 0.0))
        self.success = ok
        sel = np.nonzero(ok)[0]
        energy = 0.0
        if len(sel):
            val = O.interpolate_linear(self.tgt.image, tp[sel, 0, 0], tp[sel, 0, 1])
            self.t_patch[sel], self.dI_u[sel], self.dI_v[sel] = val[:, 0], val[:, 1], val[:, 2]
            r = (val[:, 0] - ab_t[1]) - s * (self.patch[sel] - self.ref.ab0[1])
            nrm = np.abs(r)  # PatternSize = 1
            lin = nrm * nrm > self.sigma * self.sigma
            energy = float(np.sum(np.where(lin, self.sigma * nrm - self.sigma * self.sigma / 2, nrm * nrm / 2)))
        energy += float(np.dot(ab_t * self.ab_reg, ab_t) / 2)  # AffineBrightnessPrior::energyTerm, state_priors.hpp:88-91
        # MotionPrior<SE3>::energyTerm is identically zero (state_priors.hpp:30-73)
        return energy, int(len(sel))

    def linearize(self):  # :110-192
        s, ab_t = self._scale()
        rp = O.Reprojector(self.ref, self.tgt, self.T)
        _, _, _, _, du_t, dv_t = rp.jacobians(self.uv, self.idepth)  # kCheckSuccess = false, :130
        sel = np.nonzero(self.success)[0]
        H = np.zeros((NUM_PARAMETERS, NUM_PARAMETERS))
        b = np.zeros(NUM_PARAMETERS)
        if len(sel):
            right = s * (self.patch[sel] - self.ref.ab0[1])
            r = (self.t_patch[sel] - ab_t[1]) - right
            w = np.where(r * r > self.sigma * self.sigma, self.sigma / np.maximum(np.abs(r), 1e-300), 1.0)
            d = np.zeros((len(sel), NUM_PARAMETERS))
            # leftLogTransformer of SE3 is the identity (se3_motion.hpp:239); :156-162
            d[:, :6] = -(self.dI_u[sel, None] * du_t[sel, 0, :] + self.dI_v[sel, None] * dv_t[sel, 0, :])
            d[:, 6] = -right  # :164-169
            d[:, 7] = -1.0    # :126-128
            H = (d * w[:, None]).T @ d
            b = (d * w[:, None]).T @ r
        H[6:, 6:] += np.diag(self.ab_reg)  # AffineBrightnessPrior::priorSystem, :181-185
        b[6:] += self.ab_reg * ab_t
        self.H, self.b = H, b

    def calculate_step(self, lam):  # :194-206
        H = self.H + np.diag(np.diag(self.H) * lam)
        self.step = O.normal_solve(H, self.b)
        self.old_T, self.old_ab_eps = self.T.copy(), self.ab_eps.copy()
        self.T = O.se3_exp(self.step[:6]) @ self.T  # leftIncrement, se3_motion.hpp:231-236
        self.ab_eps = self.ab_eps - self.step[6:]
        return self.step

    def accept_step(self):  # :208-213
        a = self.tgt.ab0 + self.old_ab_eps
        return float(a @ a), float(self.step @ self.step)

    def reject_step(self):  # :215-218
        self.T, self.ab_eps = self.old_T.copy(), self.old_ab_eps.copy()

    def stop(self):
        return False


def default_options():
    """createPoseAlignment, tracker/tracker/src/fabric.cpp:127-147 + EigenPoseAlignment::solve :298-305:
    lambda0 = 1 / 1e2, tolerances 1e-5, <= 50 iterations, x2 / /2, no forced accepts."""
    return O.LMOptions(50, 1e-2, 1e-5, 1e-5, False, 0, 2.0, 2.0)


def solve(ref: PAFrame, tgt: PAFrame, uv, idepth, patch, sigma_huber=20.0, ab_reg=(1e12, 1e8), opt=None, prior_rotation=None,
          trace=None):
    """EigenPoseAlignment::solve, :275-329 -> dict(rmse, energy, n_valid, converged, T_t_r, ab_eps, H, T_w_target)."""
    opt = opt or default_options()
    T = O.se3_inv(tgt.T_lin) @ ref.T_lin  # :307-308
    if prior_rotation is not None:
        T = T.copy()
        T[:3, :3] = prior_rotation  # :309-311
    p = PoseAlignerProblem(ref, tgt, uv, idepth, patch, sigma_huber, ab_reg, T)
    energy, n, conv = O.lm_solve(p, opt, trace)
    rmse = float(np.sqrt(energy / n / 1)) if n > 0 else float("inf")  # :328, PatternSize = 1
    return dict(rmse=rmse, energy=energy, n_valid=n, converged=conv, T_t_r=p.T, ab_eps=p.ab_eps, H=p.H,
                T_w_target=ref.T_lin @ O.se3_inv(p.T), iterations=len(trace) if trace is not None else None)


def solve_cpp(ref: PAFrame, tgt: PAFrame, uv, idepth, patch, sigma_huber=20.0, ab_reg=(1e12, 1e8), opt=None):
    """The same solve through the serial C++ restatement (oracle/cpu_ref/pose_alignment_cpu_ref.cpp): a second checker
    and the timed CPU baseline of the aligner.  Returns the dict of solve() plus `seconds`."""
    import ctypes as C
    import time

    from . import build_oracle
    lib = C.CDLL(build_oracle.build_pose_alignment())
    lib.paref_solve.restype = C.c_double
    opt = opt or default_options()
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rT, tT = f64(ref.T_lin[:3, :4]).reshape(12), f64(tgt.T_lin[:3, :4]).reshape(12)
    rab, tab, ri, ti = f64(ref.ab0), f64(tgt.ab0), f64(ref.intr), f64(tgt.intr)
    img = np.ascontiguousarray(tgt.image, dtype=np.float32)
    mask = np.ascontiguousarray(tgt.mask, dtype=np.uint8)
    xy, idp, pt, reg = f64(uv).reshape(-1, 2), f64(idepth), f64(patch), f64(ab_reg)
    T, ab, H = np.zeros(12), np.zeros(2), np.zeros(64)
    nv, it, cv = C.c_int(), C.c_int(), C.c_int()
    t0 = time.perf_counter()
    e = lib.paref_solve(ptr(rT), C.c_double(ref.exposure), ptr(rab), ptr(ri), C.c_int(ref.W), C.c_int(ref.H), ptr(tT),
                        C.c_double(tgt.exposure), ptr(tab), ptr(ti), C.c_int(tgt.W), C.c_int(tgt.H), ptr(img),
                        ptr(mask) if mask.min() == 0 else None, C.c_int(len(idp)), ptr(xy), ptr(idp), ptr(pt),
                        C.c_double(sigma_huber), ptr(reg), C.c_int(opt.max_num_iterations), C.c_double(opt.initial_lambda),
                        C.c_double(opt.function_tolerance), C.c_double(opt.parameter_tolerance),
                        C.c_double(opt.decrease_on_accept), C.c_double(opt.increase_on_reject), ptr(T), ptr(ab), ptr(H),
                        C.byref(nv), C.byref(it), C.byref(cv))
    dt = time.perf_counter() - t0
    T44 = np.vstack([T.reshape(3, 4), [0, 0, 0, 1.0]])
    return dict(energy=e, n_valid=nv.value, iterations=it.value, converged=bool(cv.value), T_t_r=T44, ab_eps=ab,
                H=H.reshape(8, 8), rmse=float(np.sqrt(e / nv.value)) if nv.value else float("inf"), seconds=dt)
