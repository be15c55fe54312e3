"""CPU oracle (NumPy, float64) of the immature-landmark activation refine.

TEST INFRASTRUCTURE ONLY.  PARITY PINNED: lines 122-316 of the reference file below compile here from their own source
(oracle/build_ref_tracker.py: the class, optimizeImmatureLandmark, the reference's LM driver, reprojector, PixelMap,
CameraMask; only the track containers are stand-in records) and this restatement takes the same activate / delete decision
on every candidate and the same inverse depth at 1e-9 (tests/test_reference_tracker.py, tests/golden/ref_tracker.npz);
property tests in tests/test_activation_oracle.py.

Restates src/tracker/landmarks_activator/src/landmarks_activator.cpp (paths relative to /root/reference/):
  LandmarkActivationProblem      :122-283   1-D Levenberg-Marquardt on the inverse depth of ONE immature landmark over all
                                            active frames, 8-pixel pattern, Huber weight, inlier energy cap 8 * 12^2
  optimizeImmatureLandmark       :285-316   lambda0 = 0.1, ftol = 0, ptol = 1e-8, <= 3 iterations, /2 on accept, x5 on reject;
                                            delete when valid residuals < minimum_inliers or idepth < 0
driven by energy::levenberg_marquardt_algorithm::solve (levenberg_marquardt_algorithm.hpp:77-128).
"""
from __future__ import annotations

import numpy as np

from . import pba_oracle as O
from .pose_alignment_oracle import mask_valid_checked

K_MAX_ENERGY_FOR_INLIERS = O.P * 12.0 * 12.0  # :124


class ActFrame:
    """What the problem reads from track::ActiveKeyframe: tWorldAgent, exposure, affine brightness, level-0 image, mask."""

    def __init__(self, frame_id, T_w_agent, exposure, ab, intr, image, mask):
        self.id = frame_id
        self.T_lin = np.array(T_w_agent, dtype=np.float64)
        self.exposure = float(exposure)
        self.ab = np.array(ab, dtype=np.float64)
        self.intr = np.array(intr, dtype=np.float64)
        self.image = np.asarray(image, dtype=np.float64)
        self.H, self.W = self.image.shape[:2]
        self.mask = np.asarray(mask)


class LandmarkActivationProblem:
    def __init__(self, ref: ActFrame, frames, projection, patch, sigma, idepth):
        self.ref, self.frames = ref, frames
        self.pattern = (np.asarray(projection, dtype=np.float64)[None, :] + O.PATTERN)[None, :, :]  # shiftPattern, (1, 8, 2)
        self.patch = np.asarray(patch, dtype=np.float64)
        self.sigma = float(sigma)
        self.idepth = float(idepth)
        self.old_idepth = float(idepth)
        self.hessian = 0.0
        self.b = 0.0
        self.step = 0.0
        self.stop_ = False

    def _targets(self):
        for tgt in self.frames:
            if tgt.id == self.ref.id:
                continue
            s = (tgt.exposure / self.ref.exposure) * np.exp(tgt.ab[0] - self.ref.ab[0])
            T = O.se3_inv(tgt.T_lin) @ self.ref.T_lin
            yield tgt, s, O.Reprojector(self.ref, tgt, T)

    def calculate_energy(self):  # :147-198
        if self.stop_:
            self.idepth = -1.0
            return 0.0, 0
        energy, n = 0.0, 0
        rho = np.array([self.idepth])
        for tgt, s, rp in self._targets():
            tp, ok = rp.values(self.pattern, rho)
            if not (ok[0] and mask_valid_checked(tgt.mask, tp)[0]):
                continue
            I = O.interpolate_linear(tgt.image, tp[0, :, 0], tp[0, :, 1])[:, 0]
            r = (I - tgt.ab[1]) - s * (self.patch - self.ref.ab[1])
            nrm = np.linalg.norm(r)
            w = self.sigma / nrm if nrm > self.sigma else 1.0
            if r @ r < K_MAX_ENERGY_FOR_INLIERS:
                energy += w * (r @ r)
                n += 1
            else:
                energy += K_MAX_ENERGY_FOR_INLIERS
        if n == 0:
            self.idepth = -1.0
            self.stop_ = True
        return energy, n

    def linearize(self):  # :200-250
        self.hessian, self.b = 0.0, 0.0
        rho = np.array([self.idepth])
        for tgt, s, rp in self._targets():
            tp, ok, du_id, dv_id, _, _ = rp.jacobians(self.pattern, rho)
            if not (ok[0] and mask_valid_checked(tgt.mask, np.where(np.isfinite(tp), tp, -1.0))[0]):
                continue
            val = O.interpolate_linear(tgt.image, tp[0, :, 0], tp[0, :, 1])
            r = (val[:, 0] - tgt.ab[1]) - s * (self.patch - self.ref.ab[1])
            nrm = np.linalg.norm(r)
            w = self.sigma / nrm if nrm > self.sigma else 1.0
            d = val[:, 1] * du_id[0] + val[:, 2] * dv_id[0]
            self.hessian += w * (d @ d)
            self.b += w * (d @ r)
        if self.hessian == 0:
            self.stop_ = True

    def calculate_step(self, lam):  # :252-256
        with np.errstate(divide="ignore", invalid="ignore"):
            self.step = np.float64(self.b) / np.float64(self.hessian + self.hessian * lam)
        self.old_idepth = self.idepth
        self.idepth -= self.step
        return np.array([self.step])

    def accept_step(self):  # :258
        return self.idepth * self.idepth, self.step * self.step

    def reject_step(self):  # :260
        self.idepth = self.old_idepth

    def stop(self):
        return self.stop_


def optimize_immature_landmark(ref, frames, projection, patch, idepth, minimum_inliers, sigma, trace=None):
    """optimizeImmatureLandmark :285-316 -> (activate: bool, idepth, number_of_valid_residuals)."""
    opt = O.LMOptions(3, 0.1, 0.0, 1e-8, False, 0, 2.0, 5.0)
    p = LandmarkActivationProblem(ref, frames, projection, patch, sigma, idepth)
    _, n, _ = O.lm_solve(p, opt, trace)
    activate = not (n < minimum_inliers or p.idepth < 0)
    return activate, p.idepth, n
