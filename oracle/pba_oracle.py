"""CPU oracle (NumPy, float64) for DSOPP's photometric bundle-adjustment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under dsopp_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use oracle/.

PARITY PINNED against the reference itself (RoadlyInc/DSOPP @ a4af2aa): the reference's own bundle adjustment --
LocalFrame, PixelMap, the pinhole ArrayReprojector, evaluateJacobians, firstEstimateJacobians_, the Hessian block
evaluation, the Problem class with its priors, updateMarginalizedLinearSystem, the LM driver, NormalLinearSystem --
is compiled from its sources, where they lie, against minimal stand-ins of the absent third-party libraries
(oracle/build_ref_pba.py), and every array it leaves behind on ten scenarios equals what this file computes: flags and
statuses exactly, values to 1e-9 of the array maximum (tests/test_reference_pba.py; golden vectors of the reference's
outputs in tests/golden/ref_pba.npz).  updatePointStatuses and relinearizeSystem (photometric_bundle_adjustment.cpp:307-406)
are part of that build and of the comparison (quantile threshold through the statuses it produces, outlier resets, inlier
counts, relative baselines).  Independent of that pin the
reference's own *property* tests are re-expressed in tests/test_oracle_properties.py:
  test_linear_system.cpp (J^T J, Schur and marginalisation identities),
  test_reprojects.cpp (left-perturbation reprojection Jacobians),
  test_analytical_diff.cpp (analytic vs numeric residual Jacobians),
  test_dxdy_accelerated.cpp (gradient definition).
lm_solve below additionally reproduces the reference's LM driver call by call on scripted problems
(oracle/build_ref.py, tests/test_reference_parts.py, tests/golden/ref_parts.npz).

Third-party arithmetic restated from its published closed forms (sources not under
/root/reference): Sophus @593db475 (SE3::exp, Adj, inverse; tangent = [upsilon; omega]) and
Eigen @1f4c0311 (ldlt().solve, completeOrthogonalDecomposition().pseudoInverse, JacobiSVD).

All `file:line` citations are relative to /root/reference/src/.  "PBA/" abbreviates
energy/problems/internal/energy/problems/photometric_bundle_adjustment/.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Tuple

import numpy as np

K_OK, K_OUTLIER, K_OCCLUDED, K_OOB, K_UNKNOWN = 0, 1, 2, 3, 4
PATTERN = np.array(  # common/pattern/include/common/pattern/pattern.hpp:22-33
    [[0, 2], [-1, 1], [1, 1], [-2, 0], [0, 0], [2, 0], [-1, -1], [0, -2]], dtype=np.float64
)
P = 8  # PatternSize
BLOCK = 8  # Motion::DoF + 2, PBA/eigen_photometric_bundle_adjustment_problem.hpp:259
BORDER = 4.0  # energy/camera_model/include/energy/camera_model/camera_model_base.hpp:34


# --------------------------------------------------------------------------------------------
# SE3 (Sophus restated; call sites energy/motion/include/energy/motion/se3_motion.hpp:58-252)
# --------------------------------------------------------------------------------------------
def hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def se3_exp(xi):
    xi = np.asarray(xi, dtype=np.float64)
    v, w = xi[:3], xi[3:]
    th2 = float(w @ w)
    th = np.sqrt(th2)
    W = hat(w)
    if th < 1e-10:
        R = np.eye(3) + W + 0.5 * W @ W
        V = np.eye(3) + 0.5 * W + (1.0 / 6.0) * W @ W
    else:
        R = np.eye(3) + (np.sin(th) / th) * W + ((1.0 - np.cos(th)) / th2) * W @ W
        V = np.eye(3) + ((1.0 - np.cos(th)) / th2) * W + ((th - np.sin(th)) / (th2 * th)) * W @ W
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ v
    return T


def se3_inv(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def se3_adj(T):
    """Sophus Adj = [[R, hat(t) R], [0, R]]  (se3_motion.hpp:245 rightLogTransformer)."""
    R, t = T[:3, :3], T[:3, 3]
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[:3, 3:] = hat(t) @ R
    A[3:, 3:] = R
    return A


# --------------------------------------------------------------------------------------------
# data model (PBA/local_frame.hpp:173-584)
# --------------------------------------------------------------------------------------------
class Residuals:
    """Vector of ResidualPoint for one (reference frame, target frame) -- local_frame.hpp:173-220."""

    def __init__(self, statuses: np.ndarray):
        m = len(statuses)
        self.status = np.array(statuses, dtype=np.uint8)
        self.cand = self.status.copy()
        self.r = np.zeros((m, P))
        self.du_id = np.zeros((m, P))
        self.dv_id = np.zeros((m, P))
        self.du_t = np.zeros((m, P, 6))
        self.dv_t = np.zeros((m, P, 6))
        self.jac_valid = np.zeros(m, dtype=bool)
        self.J_ref = np.zeros((m, P, BLOCK))
        self.J_tgt = np.zeros((m, P, BLOCK))
        self.d_idepth = np.zeros((m, P))
        self.w = np.ones(m)
        self.e = np.zeros(m)
        self.bcs = np.zeros(m)

    def reset(self, idx, status):
        """`residual = {status}` -- photometric_bundle_adjustment.cpp:389-391 (quirk Q6)."""
        self.status[idx] = status
        self.cand[idx] = status
        for a in (self.r, self.du_id, self.dv_id, self.du_t, self.dv_t, self.J_ref, self.J_tgt, self.d_idepth):
            a[idx] = 0
        self.jac_valid[idx] = False
        self.w[idx] = 1
        self.e[idx] = 0
        self.bcs[idx] = 0


class Frame:
    """LocalFrame -- local_frame.hpp:232-584 (single sensor, C = 1, pinhole, SE3)."""

    def __init__(self, frame_id, timestamp, T_w_lin, exposure, ab0, intr, image, mask, fixed,
                 uv, idepth, patch, flags=None, state_eps=None):
        self.id = int(frame_id)
        self.timestamp = timestamp
        self.T_lin = np.array(T_w_lin, dtype=np.float64)
        self.exposure = float(exposure)
        self.ab0 = np.array(ab0, dtype=np.float64)
        self.intr = np.array(intr, dtype=np.float64)
        self.image = np.asarray(image, dtype=np.float64)
        self.H, self.W = self.image.shape[:2]
        self.mask = np.asarray(mask)
        self.fixed = bool(fixed)
        self.is_marginalized = False
        self.to_marginalize = False
        self.state_eps = np.zeros(BLOCK) if state_eps is None else np.array(state_eps, dtype=np.float64)
        self.state_eps_step = np.zeros(BLOCK)
        m = len(idepth)
        self.uv = np.array(uv, dtype=np.float64).reshape(m, 2)
        self.idepth = np.array(idepth, dtype=np.float64)
        self.idepth_step = np.zeros(m)
        self.patch = np.array(patch, dtype=np.float64).reshape(m, P)
        flags = np.zeros(m, dtype=np.uint8) if flags is None else np.asarray(flags)
        self.lm_marginalized = (flags & 1) != 0
        self.lm_to_marginalize = (flags & 2) != 0
        self.lm_outlier = (flags & 4) != 0
        self.ill = np.zeros(m, dtype=bool)
        self.corrected = np.zeros((m, P))
        self.rel_baseline = np.zeros(m)
        self.inv_hdd = np.zeros(m)
        self.b_d = np.zeros(m)
        self.Hpd = np.zeros((m, 0))
        self.n_inliers = np.zeros(m, dtype=np.int64)
        self.residuals: Dict[int, Residuals] = {}
        self.cov: Dict[int, np.ndarray] = {}

    @property
    def ref_pattern(self):
        """reference_pattern = pattern + projection, (M, 8, 2) -- local_frame.hpp:261-266."""
        return self.uv[:, None, :] + PATTERN[None, :, :]

    def t_world_agent(self):
        """local_frame.hpp:525-527."""
        return self.T_lin @ se3_exp(self.state_eps[:6])

    def affine_brightness(self):
        return self.ab0 + self.state_eps[6:]


def frames_from_window(win, dtype_image=np.float64) -> List[Frame]:
    """Build oracle frames from a dsopp_b200.synth.SynthWindow (all-pairs connections)."""
    frames = []
    for f in win.frames:
        fr = Frame(f.frame_id, f.timestamp, f.T_w_lin, f.exposure, f.ab0, f.intr, f.image, f.mask, f.fixed,
                   f.uv, f.idepth, f.patch, f.flags, f.state_eps)
        frames.append(fr)
    for (r, t), st in win.statuses.items():
        frames[r].residuals[frames[t].id] = Residuals(st)
    return frames


# --------------------------------------------------------------------------------------------
# camera / sampler
# --------------------------------------------------------------------------------------------
def inside_roi(pts, W, H):
    """camera_model_base.hpp:52-60 on (..., 8, 2) patterns -> (...,) bool (all 8 points)."""
    x, y = pts[..., 0], pts[..., 1]
    ok = (x >= BORDER) & (y >= BORDER) & (x <= W - BORDER - 1) & (y <= H - BORDER - 1)
    return ok.all(axis=-1)


def valid_idepth(rho):
    """camera_model_base.hpp:68-74: -1e-4 < rho < 1/0.001 + 10."""
    return (rho > -1e-4) & (rho < 1.0 / 0.001 + 1e1)


def mask_valid(mask, pts):
    """CameraMask::valid<false>: round() then uchar lookup, camera_mask.hpp:48-89 (quirk Q5)."""
    xi = np.floor(np.abs(pts[..., 0]) + 0.5).astype(np.int64) * np.sign(pts[..., 0]).astype(np.int64)
    yi = np.floor(np.abs(pts[..., 1]) + 0.5).astype(np.int64) * np.sign(pts[..., 1]).astype(np.int64)
    return (mask[yi, xi] != 0).all(axis=-1)


def interpolate_linear(image, x, y):
    """features/include/features/camera/pixel_map.hpp:20-40: truncation + 4 weighted taps of {I,dx,dy}."""
    ix = x.astype(np.int64)  # static_cast<int> truncation
    iy = y.astype(np.int64)
    dx = x - ix
    dy = y - iy
    dxdy = dx * dy
    return (dxdy[..., None] * image[iy + 1, ix + 1]
            + (dy - dxdy)[..., None] * image[iy + 1, ix]
            + (dx - dxdy)[..., None] * image[iy, ix + 1]
            + (1 - dx - dy + dxdy)[..., None] * image[iy, ix])


class Reprojector:
    """ArrayReprojector<T, PinholeCamera, SE3> -- energy/projector/include/energy/projector/camera_reproject.hpp:194-382."""

    def __init__(self, ref: Frame, tgt: Frame, T_t_r):
        fxr, fyr, cxr, cyr = ref.intr
        fxt, fyt, cxt, cyt = tgt.intr
        K = np.array([[fxt, 0, cxt], [0, fyt, cyt], [0, 0, 1.0]])
        Kinv = np.eye(4)
        Kinv[0, 0] = 1 / fxr
        Kinv[1, 1] = 1 / fyr
        Kinv[0, 2] = -cxr / fxr
        Kinv[1, 2] = -cyr / fyr
        self.reproject_ = K @ T_t_r[:3, :4] @ Kinv  # :256
        self.project_ = K
        self.transform_unproject_ = T_t_r[:3, :4] @ Kinv  # :258
        self.translation_ = T_t_r[:3, 3].copy()
        self.ref, self.tgt = ref, tgt

    def values(self, ref_pts, rho):
        """:270-293.  ref_pts (M,8,2), rho (M,) -> target (M,8,2), success (M,)."""
        A = self.reproject_
        p = ref_pts @ A[:, :2].T + (A[:, 2][None, None, :] + rho[:, None, None] * A[:, 3][None, None, :])
        with np.errstate(divide="ignore", invalid="ignore"):
            tgt = p[..., :2] / p[..., 2:3]
        ok = valid_idepth(rho) & inside_roi(ref_pts, self.ref.W, self.ref.H)
        ok &= (p[..., 2] > 0).all(axis=-1)
        with np.errstate(invalid="ignore"):
            ok &= inside_roi(np.where(np.isfinite(tgt), tgt, -1.0), self.tgt.W, self.tgt.H)
        return tgt, ok

    def jacobians(self, ref_pts, rho):
        """:305-367 -> target, success, du_idepth, dv_idepth (M,8), du_t, dv_t (M,8,6)  (quirk Q4)."""
        Mx = self.transform_unproject_
        q = ref_pts @ Mx[:, :2].T + (Mx[:, 2][None, None, :] + rho[:, None, None] * Mx[:, 3][None, None, :])
        ok = valid_idepth(rho) & inside_roi(ref_pts, self.ref.W, self.ref.H)
        ok &= (q[..., 2] > 0).all(axis=-1)
        proj = q @ self.project_.T
        with np.errstate(divide="ignore", invalid="ignore"):
            tgt = proj[..., :2] / proj[..., 2:3]
            ok &= inside_roi(np.where(np.isfinite(tgt), tgt, -1.0), self.tgt.W, self.tgt.H)
            s = 1.0 / q[..., 2]
            b0 = q[..., 0] * s
            b1 = q[..., 1] * s
            fx, fy = self.project_[0, 0], self.project_[1, 1]
            t = self.translation_
            du_id = fx * (t[0] * s - t[2] * s * b0)
            dv_id = fy * (t[1] * s - t[2] * s * b1)
            nid = rho[:, None] * s
            z = np.zeros_like(b0)
            dv_t = fy * np.stack([z, nid, -nid * b1, -(b1 * b1 + 1), b0 * b1, b0], axis=-1)
            du_t = fx * np.stack([nid, z, -nid * b0, -(b0 * b1), b0 * b0 + 1, -b1], axis=-1)
        return tgt, ok, du_id, dv_id, du_t, dv_t


def relative_pose(ref: Frame, tgt: Frame, with_step=True):
    """PBA/evaluate_jacobians.hpp:36-49: T_t_r0 and T_t_r = exp(-eps_t) T_t_r0 exp(eps_r)."""
    T0 = se3_inv(tgt.T_lin) @ ref.T_lin
    er = ref.state_eps[:6] + (ref.state_eps_step[:6] if with_step else 0)
    et = tgt.state_eps[:6] + (tgt.state_eps_step[:6] if with_step else 0)
    return T0, se3_exp(-et) @ (T0 @ se3_exp(er))


# --------------------------------------------------------------------------------------------
# K6: firstEstimateJacobians_  (PBA/first_estimate_jacobians.hpp:14-71)
# --------------------------------------------------------------------------------------------
def first_estimate_jacobians(frames: List[Frame]):
    for ref in frames:
        for tgt in frames:  # deque order; the last target wins corrected_intensities (quirk Q1)
            if ref.id == tgt.id or tgt.id not in ref.residuals:
                continue
            T0 = se3_inv(tgt.T_lin) @ ref.T_lin
            rp = Reprojector(ref, tgt, T0)
            s0 = (tgt.exposure / ref.exposure) * np.exp(tgt.ab0[0] - ref.ab0[0])
            res = ref.residuals[tgt.id]
            sel = ~(ref.lm_marginalized & ~ref.lm_to_marginalize)
            idx = np.nonzero(sel)[0]
            if len(idx) == 0:
                continue
            _, ok, du_id, dv_id, du_t, dv_t = rp.jacobians(ref.ref_pattern[idx], ref.idepth[idx])  # Q9: current idepth
            res.jac_valid[idx] = ok
            res.du_id[idx], res.dv_id[idx], res.du_t[idx], res.dv_t[idx] = du_id, dv_id, du_t, dv_t
            ref.corrected[idx] = s0 * (ref.patch[idx] - ref.ab0[1])
            res.bcs[idx] = s0


# --------------------------------------------------------------------------------------------
# K1/K2: evaluateJacobians  (PBA/evaluate_jacobians.hpp:20-202)
# --------------------------------------------------------------------------------------------
def evaluate_jacobians(frames: List[Frame], sigma_huber: float = 0.0, *, fej: bool, evaluate_jacobians: bool,
                       new_point: bool = True, huber: bool = False, optimize_idepths: bool = True):
    assert huber or sigma_huber == 0
    sig2 = sigma_huber * sigma_huber
    for ref in frames:
        for tgt in frames:
            if ref.id == tgt.id or tgt.id not in ref.residuals:
                continue
            T0, T = relative_pose(ref, tgt)
            ab_r = ref.ab0 + ref.state_eps[6:] + ref.state_eps_step[6:]
            ab_t = tgt.ab0 + tgt.state_eps[6:] + tgt.state_eps_step[6:]
            s = (tgt.exposure / ref.exposure) * np.exp(ab_t[0] - ab_r[0])  # :56-57
            rp = Reprojector(ref, tgt, T)
            right_log = se3_adj(T0 if fej else T)  # :62-64
            left_log = np.eye(6)  # :65-66 (identity for SE3)
            res = ref.residuals[tgt.id]
            sel = ~(ref.lm_marginalized & ~ref.lm_to_marginalize)  # :83
            idx = np.nonzero(sel)[0]
            if len(idx) == 0:
                continue
            pat = ref.ref_pattern[idx]
            rho = ref.idepth[idx] + ref.idepth_step[idx]
            if fej or not evaluate_jacobians:  # :91-96
                tp, ok = rp.values(pat, rho)
                ok = ok & ((not fej) | res.jac_valid[idx])
                dshift = res.bcs[idx].copy()
                corrected = ref.corrected[idx].copy()
            else:  # :97-108
                tp, ok, du_id, dv_id, du_t, dv_t = rp.jacobians(pat, rho)
                res.jac_valid[idx] = ok
                res.du_id[idx], res.dv_id[idx], res.du_t[idx], res.dv_t[idx] = du_id, dv_id, du_t, dv_t
                corrected = s * (ref.patch[idx] - ab_r[1])
                dshift = np.full(len(idx), s)
            okm = ok.copy()
            if ok.any():  # mask lookup only where the ROI test passed (Q5)
                okm[ok] = mask_valid(tgt.mask, tp[ok])
            ok = okm
            res.cand[idx[~ok]] = K_OOB  # :111-113
            ev = ok & (res.status[idx] == K_OK)  # :114
            ie = idx[ev]
            ine = idx[~ev]
            if len(ie):
                res.cand[ie] = K_OK
                tpe = tp[ev]
                samp = interpolate_linear(tgt.image, tpe[..., 0], tpe[..., 1])  # (m,8,3)
                left = samp[..., 0] - ab_t[1]  # :124-126
                right = s * (ref.patch[ie] - ab_r[1])  # :127-130
                if new_point:
                    r = left - right  # measures/include/measures/similarity_measure_ssd.hpp:30-33
                    res.r[ie] = r
                    n2 = (r * r).sum(axis=1)
                    e = 0.5 * n2
                    w = np.ones(len(ie))
                    if huber:
                        big = n2 > sig2
                        nrm = np.sqrt(n2[big])
                        w[big] = sigma_huber / nrm
                        e[big] = sigma_huber * nrm - sig2 * 0.5
                    res.e[ie] = e
                    res.w[ie] = w
                if evaluate_jacobians:
                    dIu = samp[..., 1]
                    dIv = samp[..., 2]
                    Jg = dIv[..., None] * res.dv_t[ie] + dIu[..., None] * res.du_t[ie]  # :149-157
                    res.J_tgt[ie, :, :6] = -(Jg @ left_log)  # :159-160
                    res.J_ref[ie, :, :6] = Jg @ right_log  # :162-163
                    if optimize_idepths:
                        res.d_idepth[ie] = dIu * res.du_id[ie] + dIv * res.dv_id[ie]  # :165-174
                    c = corrected[ev]
                    res.J_ref[ie, :, 6] = c  # :178
                    res.J_ref[ie, :, 7] = dshift[ev][:, None]  # :179
                    res.J_tgt[ie, :, 6] = -c  # :181
                    res.J_tgt[ie, :, 7] = -1.0  # :182
            if len(ine):  # :184-194
                if new_point:
                    res.r[ine] = 0
                    res.e[ine] = 0
                if evaluate_jacobians:
                    res.J_ref[ine] = 0
                    res.J_tgt[ine] = 0
                    res.d_idepth[ine] = 0


def change_residual_statuses(frames: List[Frame], accept: bool = True):
    """PBA/eigen_photometric_bundle_adjustment_problem.hpp:20-35."""
    for f in frames:
        for res in f.residuals.values():
            if accept:
                res.status[:] = res.cand
            else:
                res.cand[:] = res.status


# --------------------------------------------------------------------------------------------
# K3: evaluateLinearSystemPosePose  (PBA/hessian_block_evaluation.hpp:38-164)
# --------------------------------------------------------------------------------------------
def _lm_select(ref: Frame, for_marginalized: bool):
    return ref.lm_to_marginalize if for_marginalized else ~ref.lm_marginalized  # :68-72


def pose_pose_block(ref: Frame, tgt: Frame, for_marginalized=False):
    res = ref.residuals[tgt.id]
    sel = _lm_select(ref, for_marginalized)
    w = res.w[sel]
    Jr, Jt, r = res.J_ref[sel], res.J_tgt[sel], res.r[sel]
    Hrr = np.einsum("l,lpi,lpj->ij", w, Jr, Jr)
    Htt = np.einsum("l,lpi,lpj->ij", w, Jt, Jt)
    Hrt = np.einsum("l,lpi,lpj->ij", w, Jr, Jt)
    br = np.einsum("l,lpi,lp->i", w, Jr, r)
    bt = np.einsum("l,lpi,lp->i", w, Jt, r)
    return Hrr, Hrt, Htt, br, bt


def pose_pose(frames: List[Frame], for_marginalized=False) -> Tuple[np.ndarray, np.ndarray]:
    n = len(frames)
    H = np.zeros((BLOCK * n, BLOCK * n))
    b = np.zeros(BLOCK * n)
    for ri, ref in enumerate(frames):
        for ti, tgt in enumerate(frames):
            if ri == ti:
                continue
            assert tgt.id in ref.residuals, "window must be fully connected (quirk Q2)"
            Hrr, Hrt, Htt, br, bt = pose_pose_block(ref, tgt, for_marginalized)
            R, T = slice(BLOCK * ri, BLOCK * ri + BLOCK), slice(BLOCK * ti, BLOCK * ti + BLOCK)
            H[R, R] += Hrr
            H[R, T] = Hrt  # assignment, not += (quirk Q3, :125-128)
            H[T, T] += Htt
            b[R] += br
            b[T] += bt
        assert (not ref.fixed) or ri == 0  # :143
    # symmetrise (:147-163)
    for i in range(n):
        I = slice(BLOCK * i, BLOCK * i + BLOCK)
        D = H[I, I]
        H[I, I] = np.tril(D) + np.tril(D, -1).T  # selfadjointView<Lower>
        for j in range(i + 1, n):
            J = slice(BLOCK * j, BLOCK * j + BLOCK)
            H[I, J] += H[J, I].T
            H[J, I] = H[I, J].T
    return H, b


# --------------------------------------------------------------------------------------------
# K4: evaluateLinearSystemPoseDepthSchurComplement  (hessian_block_evaluation.hpp:169-236)
# --------------------------------------------------------------------------------------------
def schur_complement(frames: List[Frame], for_marginalized=False) -> Tuple[np.ndarray, np.ndarray]:
    n = len(frames)
    Hs = np.zeros((BLOCK * n, BLOCK * n))
    bs = np.zeros(BLOCK * n)
    for ri, ref in enumerate(frames):
        m = len(ref.idepth)
        sel = _lm_select(ref, for_marginalized)
        Hpd = np.zeros((m, BLOCK * n))
        hdd = np.zeros(m)
        bd = np.zeros(m)
        for ti, tgt in enumerate(frames):
            if ri == ti:
                continue
            res = ref.residuals[tgt.id]
            Hpd[:, BLOCK * ri:BLOCK * ri + BLOCK] += np.einsum("l,lpi,lp->li", res.w, res.J_ref, res.d_idepth)
            Hpd[:, BLOCK * ti:BLOCK * ti + BLOCK] += np.einsum("l,lpi,lp->li", res.w, res.J_tgt, res.d_idepth)
            hdd += res.w * (res.d_idepth * res.d_idepth).sum(axis=1)
            bd += res.w * (res.d_idepth * res.r).sum(axis=1)
        if ref.Hpd.shape != (m, BLOCK * n):
            ref.Hpd = np.zeros((m, BLOCK * n))
        ref.b_d[sel] = bd[sel]  # :213-214
        ref.Hpd[sel] = Hpd[sel]
        good = sel & (hdd > 1e-15)  # :215-216
        if for_marginalized and ref.fixed:
            hdd = hdd + 1e8  # kScaleNullspaceRegularizer, :217-219
        inv = np.zeros(m)
        inv[good] = 1.0 / hdd[good]
        ref.inv_hdd[good] = inv[good]
        ref.ill[good] = False
        ref.ill[sel & ~good] = True
        bs += (inv[good] * bd[good]) @ Hpd[good]
        Hs += (Hpd[good] * inv[good][:, None]).T @ Hpd[good]
    return Hs, bs


# --------------------------------------------------------------------------------------------
# K5: calculateIdepths  (hessian_block_evaluation.hpp:238-263)
# --------------------------------------------------------------------------------------------
def calculate_idepths(frames: List[Frame], step_poses: np.ndarray, lam: float):
    k = 1.0 / (1.0 + lam)
    for ref in frames:
        sel = (~ref.lm_marginalized) & (~ref.ill)
        if ref.Hpd.shape[1] != len(step_poses):
            continue
        step = (ref.b_d[sel] - ref.Hpd[sel] @ step_poses) * k * ref.inv_hdd[sel]
        ref.idepth_step[sel] = -step


# --------------------------------------------------------------------------------------------
# priors, energy, state  (PBA/eigen_photometric_bundle_adjustment_problem.hpp:37-144, state_priors.hpp)
# --------------------------------------------------------------------------------------------
def linear_system_prior(frames, H, b, ab_reg, fixed_reg, for_marginalized=False):
    """evaluateLinearSystemPrior :37-77 (MotionPrior<SE3> is identically zero, state_priors.hpp:30-73)."""
    for i, f in enumerate(frames):
        if f.to_marginalize != for_marginalized:
            continue
        o = BLOCK * i
        if f.fixed:
            H[o:o + BLOCK, o:o + BLOCK] += np.eye(BLOCK) * fixed_reg
            b[o:o + BLOCK] += fixed_reg * f.state_eps
        else:
            ab = f.ab0 + f.state_eps[6:]
            H[o + 6:o + 8, o + 6:o + 8] += np.diag(ab_reg)
            b[o + 6:o + 8] += ab_reg * ab


def state_eps_stacked(frames, with_step=False):
    s = np.concatenate([f.state_eps for f in frames])
    if with_step:
        s = s + np.concatenate([f.state_eps_step for f in frames])
    return s


def landmarks_energy(frames, for_marginalized=False):
    """calculateLandmarksEnergy :93-144."""
    e, n = 0.0, 0
    for ref in frames:
        sel = _lm_select(ref, for_marginalized)
        for tgt in frames:
            if tgt.id == ref.id or tgt.id not in ref.residuals:
                continue
            res = ref.residuals[tgt.id]
            e += float(res.e[sel].sum())
            n += int((res.e[sel] > 0).sum())
    return e, n


# --------------------------------------------------------------------------------------------
# NormalLinearSystem  (energy/problems/src/normal_linear_system.cpp:10-59)
# --------------------------------------------------------------------------------------------
def jacobi_preconditioner(H):
    return 1.0 / np.sqrt(np.diag(H) + 10.0)  # kPreconditionerMinValue, :12-15


def normal_solve(H, b):
    """:51-59: x = p * LDLT(p H p).solve(p b).  (Eigen LDLT restated by a dense symmetric solve.)"""
    p = jacobi_preconditioner(H)
    Hp = H * p[:, None] * p[None, :]
    return p * np.linalg.solve(Hp, p * b)


def reduce_system(H, b, elim: List[int]):
    """:18-50: Schur-eliminate `elim` with Jacobi preconditioning, COD pseudo-inverse, symmetrisation."""
    n = len(b)
    elim = list(elim)
    keep = [i for i in range(n) if i not in set(elim)]
    p = jacobi_preconditioner(H)
    pinv = 1.0 / p
    Hp = H * p[:, None] * p[None, :]
    bp = p * b
    St = Hp[np.ix_(keep, elim)] @ np.linalg.pinv(Hp[np.ix_(elim, elim)])
    Hn = Hp[np.ix_(keep, keep)] - St @ Hp[np.ix_(keep, elim)].T
    bn = bp[keep] - St @ bp[elim]
    Hn = 0.5 * (Hn + Hn.T)
    pk = pinv[keep]
    return Hn * pk[:, None] * pk[None, :], pk * bn


# --------------------------------------------------------------------------------------------
# PhotometricBundleAdjustmentProblem  (eigen_photometric_bundle_adjustment_problem.hpp:255-429)
# --------------------------------------------------------------------------------------------
class Problem:
    def __init__(self, frames: List[Frame], sigma_huber: float, H_marg=None, b_marg=None, energy_marg=0.0,
                 ab_reg=(1e12, 1e8), fixed_reg=1e16, fej=True):
        self.frames = frames
        self.sigma = float(sigma_huber)
        n = BLOCK * len(frames)
        self.H_marg = np.zeros((n, n)) if H_marg is None else np.array(H_marg, dtype=np.float64)
        self.b_marg = np.zeros(n) if b_marg is None else np.array(b_marg, dtype=np.float64)
        self.energy_marg = float(energy_marg)
        self.ab_reg = np.array(ab_reg, dtype=np.float64)
        self.fixed_reg = float(fixed_reg)
        self.fej = fej
        self.H_pose = np.zeros((n, n))
        self.b_pose = np.zeros(n)
        self.H_schur = np.zeros((n, n))
        self.b_schur = np.zeros(n)

    def calculate_energy(self):  # :290-317
        evaluate_jacobians(self.frames, self.sigma, fej=self.fej, evaluate_jacobians=False, new_point=True, huber=True)
        s = state_eps_stacked(self.frames, True)
        energy = self.energy_marg + self.b_marg @ s + 0.5 * (s @ (self.H_marg @ s))  # DSO eq 8.19
        for f in self.frames:  # every frame, fixed included (quirk Q8)
            ab = f.ab0 + f.state_eps[6:] + f.state_eps_step[6:]
            energy += 0.5 * float((ab * self.ab_reg) @ ab)  # state_priors.hpp:87-90
        le, nv = landmarks_energy(self.frames)
        return energy + le, nv

    def linearize(self):  # :322-336
        evaluate_jacobians(self.frames, self.sigma, fej=self.fej, evaluate_jacobians=True, new_point=True, huber=True)
        self.H_pose, self.b_pose = pose_pose(self.frames)
        linear_system_prior(self.frames, self.H_pose, self.b_pose, self.ab_reg, self.fixed_reg)
        self.H_schur, self.b_schur = schur_complement(self.frames)

    def calculate_step(self, lam):  # :342-361
        state = state_eps_stacked(self.frames)
        H = self.H_pose + self.H_marg
        b = self.b_pose + self.b_marg
        H[np.diag_indices_from(H)] += np.diag(self.H_pose) * lam
        k = -1.0 / (1.0 + lam)
        H = H + self.H_schur * k
        b = b + self.b_schur * k
        b = b + self.H_marg @ state
        step = normal_solve(H, b)
        for i, f in enumerate(self.frames):
            f.state_eps_step = -step[BLOCK * i:BLOCK * i + BLOCK]
        calculate_idepths(self.frames, step, lam)
        return step

    def accept_step(self):  # :366-388
        state_sq, step_sq = 0.0, 0.0
        for f in self.frames:
            state_sq += float(f.state_eps @ f.state_eps) + float(f.ab0 @ f.ab0)
            f.state_eps = f.state_eps + f.state_eps_step
            step_sq += float(f.state_eps_step @ f.state_eps_step)
            f.state_eps_step = np.zeros(BLOCK)
            state_sq += float(f.idepth @ f.idepth)
            f.idepth = f.idepth + f.idepth_step
            step_sq += float(f.idepth_step @ f.idepth_step)
            f.idepth_step = np.zeros_like(f.idepth_step)
        change_residual_statuses(self.frames, True)
        return state_sq, step_sq

    def reject_step(self):  # :392-402
        for f in self.frames:
            f.state_eps_step = np.zeros(BLOCK)
            f.idepth_step = np.zeros_like(f.idepth_step)
        change_residual_statuses(self.frames, False)

    def stop(self):
        return False


# --------------------------------------------------------------------------------------------
# levenberg_marquardt_algorithm::solve
# (energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:77-128)
# --------------------------------------------------------------------------------------------
class LMOptions:
    def __init__(self, max_num_iterations=50, initial_lambda=1e-5, function_tolerance=1e-8, parameter_tolerance=1e-8,
                 force_accept=False, min_num_iterations=0, decrease_on_accept=2.0, increase_on_reject=10.0):
        self.max_num_iterations = max_num_iterations
        self.initial_lambda = initial_lambda
        self.function_tolerance = function_tolerance
        self.parameter_tolerance = parameter_tolerance
        self.force_accept = force_accept
        self.min_num_iterations = min_num_iterations
        self.decrease_on_accept = decrease_on_accept
        self.increase_on_reject = increase_on_reject


def lm_solve(problem, opt: LMOptions, trace: Optional[list] = None):
    lam = opt.initial_lambda
    energy, nvalid = problem.calculate_energy()
    converged = False
    system_valid = False
    it = 0
    while it < opt.max_num_iterations and not converged and nvalid > 0:
        if not system_valid:
            problem.linearize()
        step = problem.calculate_step(lam)
        next_energy, next_n = problem.calculate_energy()
        if problem.stop() or next_n == 0:
            problem.reject_step()
            break
        converged |= abs(energy - next_energy) / energy < opt.function_tolerance  # quirk Q7
        if next_energy < energy or (opt.force_accept and it < opt.min_num_iterations):
            state_sq, step_sq = problem.accept_step()
            converged |= step_sq < opt.parameter_tolerance * (state_sq + opt.parameter_tolerance)
            energy, nvalid = next_energy, next_n
            lam /= opt.decrease_on_accept
            system_valid = False
            if trace is not None:
                trace.append(dict(it=it, accepted=True, energy=next_energy, n=next_n, step=np.array(step)))
        else:
            problem.reject_step()
            if trace is not None:
                trace.append(dict(it=it, accepted=False, energy=next_energy, n=next_n, step=np.array(step)))
            if opt.force_accept:
                problem.calculate_energy()
                return energy, nvalid, converged
            lam *= opt.increase_on_reject
            system_valid = True
        it += 1
    problem.calculate_energy()
    return energy, nvalid, converged


# --------------------------------------------------------------------------------------------
# marginalisation  (eigen_photometric_bundle_adjustment_problem.hpp:146-203)
# --------------------------------------------------------------------------------------------
def update_marginalized_linear_system(frames: List[Frame], H_marg, b_marg, energy_marg, ab_reg, fixed_reg):
    """Returns (frames', H_marg', b_marg', energy_marg')."""
    Hs, bs = schur_complement(frames, for_marginalized=True)
    Hp, bp = pose_pose(frames, for_marginalized=True)
    H = Hp - Hs
    b = bp - bs
    state = state_eps_stacked(frames)
    le, _ = landmarks_energy(frames, for_marginalized=True)
    energy_marg = energy_marg + le + state @ (H @ state) - state @ b  # DSO eq 8.15, :169-170
    b = b - H @ state
    H_marg = H_marg + H
    b_marg = b_marg + b
    for f in frames:
        f.lm_to_marginalize[:] = False
    elim = [BLOCK * i + p for i, f in enumerate(frames) if f.to_marginalize for p in range(BLOCK)]
    if not elim:
        return frames, H_marg, b_marg, energy_marg
    n = BLOCK * len(frames)
    Hpr, bpr = np.zeros((n, n)), np.zeros(n)
    linear_system_prior(frames, Hpr, bpr, np.asarray(ab_reg, dtype=np.float64), fixed_reg, for_marginalized=True)
    bpr = bpr - Hpr @ state
    H_marg = H_marg + Hpr
    b_marg = b_marg + bpr
    H_marg, b_marg = reduce_system(H_marg, b_marg, elim)
    frames = [f for f in frames if not f.to_marginalize]
    return frames, H_marg, b_marg, energy_marg


# --------------------------------------------------------------------------------------------
# updatePointStatuses  (energy/problems/src/photometric_bundle_adjustment.cpp:322-406)
# --------------------------------------------------------------------------------------------
def update_point_statuses(frames: List[Frame], min_valid: int, sigma_huber: float):
    energies = []
    for ref in frames:
        act = ~ref.lm_marginalized
        for tgt in frames:
            if tgt.is_marginalized or tgt.id == ref.id or tgt.id not in ref.residuals:
                continue
            res = ref.residuals[tgt.id]
            energies.append(res.e[act & (res.status == K_OK)])
    energies = np.concatenate(energies) if energies else np.zeros(0)
    if len(energies):
        k = int(float(len(energies)) * 0.75)
        thr = np.partition(energies, k)[k] + sigma_huber * sigma_huber / 2  # nth_element, :358-361
    else:
        thr = 0.0
    for ref in frames:
        act = ~ref.lm_marginalized
        valid = np.zeros(len(ref.idepth), dtype=np.int64)
        ref.n_inliers[act] = 0
        for tgt in frames:
            if tgt.is_marginalized or tgt.id == ref.id or tgt.id not in ref.residuals:
                continue
            res = ref.residuals[tgt.id]
            dist = np.linalg.norm(ref.t_world_agent()[:3, 3] - tgt.t_world_agent()[:3, 3])
            out = act & (res.e > thr)
            res.reset(np.nonzero(out)[0], K_OUTLIER)
            ok = act & (res.status == K_OK)
            ref.rel_baseline[ok] = np.maximum(ref.rel_baseline[ok], ref.idepth[ok] * dist)
            valid[ok] += 1
            ref.n_inliers[ok] += 1
        ref.lm_outlier[act & (valid < min_valid)] = True
    return thr


# --------------------------------------------------------------------------------------------
# uncertainty  (eigen_photometric_bundle_adjustment.cpp:31-45, problem.hpp:204-242, covariance_...hpp)
# --------------------------------------------------------------------------------------------
def pseudo_inverse(M, n_null):
    U, D, Vt = np.linalg.svd(M)
    S = np.zeros_like(D)
    k = len(D) - n_null
    S[:k] = 1.0 / D[:k]
    return (Vt.T * S[None, :]) @ U.T


def covariance_pose_pose(frames, H_marg, ab_reg, fixed_reg, fej=True):
    evaluate_jacobians(frames, 0.0, fej=fej, evaluate_jacobians=True, new_point=True, huber=False)
    Hp, bp = pose_pose(frames)
    linear_system_prior(frames, Hp, bp, np.asarray(ab_reg, dtype=np.float64), fixed_reg)
    Hs, _ = schur_complement(frames)
    return pseudo_inverse(Hp - Hs + H_marg, 1)


def covariances_of_relative_poses(frames, cov):
    for ri, ref in enumerate(frames):
        for ti, tgt in enumerate(frames):
            if ri == ti:
                continue
            s11 = cov[BLOCK * ri:BLOCK * ri + 6, BLOCK * ri:BLOCK * ri + 6]
            s22 = cov[BLOCK * ti:BLOCK * ti + 6, BLOCK * ti:BLOCK * ti + 6]
            s12 = cov[BLOCK * ri:BLOCK * ri + 6, BLOCK * ti:BLOCK * ti + 6]
            adj = se3_adj(se3_inv(tgt.t_world_agent()) @ ref.t_world_agent())  # se3_motion.hpp:151-158
            ref.cov[tgt.id] = adj @ s11 @ adj.T - s12.T @ adj.T - adj @ s12 + s22


# --------------------------------------------------------------------------------------------
# EigenPhotometricBundleAdjustment::solve  (energy/problems/src/eigen_photometric_bundle_adjustment.cpp:59-101)
# --------------------------------------------------------------------------------------------
class EigenPBA:
    def __init__(self, max_iterations=7, trust_region_radius=1e5, function_tolerance=1e-8, parameter_tolerance=1e-8,
                 ab_reg=(1e12, 1e8), fixed_reg=1e16, sigma_huber=20.0, estimate_uncertainty=True, force_accept=True,
                 fej=True):
        self.max_iterations = max_iterations
        self.radius = trust_region_radius
        self.ftol, self.ptol = function_tolerance, parameter_tolerance
        self.ab_reg = np.array(ab_reg, dtype=np.float64)
        self.fixed_reg = fixed_reg
        self.sigma = sigma_huber
        self.estimate_uncertainty = estimate_uncertainty
        self.force_accept = force_accept
        self.fej = fej
        self.frames: List[Frame] = []
        self.H_marg = np.zeros((0, 0))
        self.b_marg = np.zeros(0)
        self.energy_marg = 0.0

    def set_frames(self, frames):
        self.frames = frames
        n = BLOCK * len(frames)
        H = np.zeros((n, n))
        b = np.zeros(n)
        k = min(n, len(self.b_marg))
        H[:k, :k] = self.H_marg[:k, :k]
        b[:k] = self.b_marg[:k]
        self.H_marg, self.b_marg = H, b

    def marginalize(self):
        """The frames_.size()>1 part of pushFrame, :121-130."""
        if self.fej:
            first_estimate_jacobians(self.frames)
        evaluate_jacobians(self.frames, self.sigma, fej=self.fej, evaluate_jacobians=True, new_point=True, huber=True)
        change_residual_statuses(self.frames)
        self.frames, self.H_marg, self.b_marg, self.energy_marg = update_marginalized_linear_system(
            self.frames, self.H_marg, self.b_marg, self.energy_marg, self.ab_reg, self.fixed_reg)

    def solve(self, trace=None):
        opt = LMOptions(self.max_iterations, 1.0 / self.radius, self.ftol, self.ptol, self.force_accept, 3, 1.0, 1.0)
        prob = Problem(self.frames, self.sigma, self.H_marg, self.b_marg, self.energy_marg, self.ab_reg,
                       self.fixed_reg, self.fej)
        if self.fej:
            first_estimate_jacobians(self.frames)
        energy, _, _ = lm_solve(prob, opt, trace)
        last = self.frames[-1]  # relinearizeSystem, photometric_bundle_adjustment.cpp:311-316 (quirk Q9)
        last.T_lin = last.t_world_agent()
        last.ab0 = last.affine_brightness()
        last.state_eps = np.zeros(BLOCK)
        if self.estimate_uncertainty:
            if self.fej:
                first_estimate_jacobians(self.frames)
            cov = covariance_pose_pose(self.frames, self.H_marg, self.ab_reg, self.fixed_reg, self.fej)
            covariances_of_relative_poses(self.frames, cov)
        update_point_statuses(self.frames, 1, self.sigma)
        return energy


def clone_frames(frames):
    return copy.deepcopy(frames)
