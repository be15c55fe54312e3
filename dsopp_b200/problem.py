"""Python mirror of the host-side LM problem on top of the C ABI (test / bench plumbing).

The shipped host side is C++ (dsopp_b200/csrc/host/, same structure); this mirror exists so that pytest can
drive the C ABI directly and cross-check the C++ adapter.  It follows
PhotometricBundleAdjustmentProblem (src/energy/problems/internal/energy/problems/photometric_bundle_adjustment/
eigen_photometric_bundle_adjustment_problem.hpp:255-429) and levenberg_marquardt_algorithm::solve
(src/energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:77-128):
the device runs the sweeps, the host keeps priors, the marginalised prior and the 8N x 8N solve in float64.
"""
from __future__ import annotations

import numpy as np

from .capi import BLOCK, Handle


def jacobi_solve(H, b):
    """NormalLinearSystem::solve (src/energy/problems/src/normal_linear_system.cpp:51-59)."""
    p = 1.0 / np.sqrt(np.diag(H) + 10.0)
    return p * np.linalg.solve(H * p[:, None] * p[None, :], p * b)


class CudaProblem:
    def __init__(self, handle: Handle, frames_meta, sigma_huber, H_marg=None, b_marg=None, energy_marg=0.0,
                 ab_reg=(1e12, 1e8), fixed_reg=1e16, fej=True):
        """frames_meta: list of dicts with keys ab0 (2,), fixed (bool) for every frame slot."""
        self.h = handle
        self.meta = frames_meta
        self.sigma = float(sigma_huber)
        n = BLOCK * len(frames_meta)
        self.H_marg = np.zeros((n, n)) if H_marg is None else np.array(H_marg, dtype=np.float64)
        self.b_marg = np.zeros(n) if b_marg is None else np.array(b_marg, dtype=np.float64)
        self.energy_marg = float(energy_marg)
        self.ab_reg = np.array(ab_reg, dtype=np.float64)
        self.fixed_reg = float(fixed_reg)
        self.fej = fej
        self.H_pose = self.b_pose = self.H_schur = self.b_schur = None

    def _prior(self, H, b, eps):
        """evaluateLinearSystemPrior, problem.hpp:37-77."""
        for i, m in enumerate(self.meta):
            o = BLOCK * i
            if m["fixed"]:
                H[o:o + BLOCK, o:o + BLOCK] += np.eye(BLOCK) * self.fixed_reg
                b[o:o + BLOCK] += self.fixed_reg * eps[o:o + BLOCK]
            else:
                ab = np.asarray(m["ab0"]) + eps[o + 6:o + 8]
                H[o + 6:o + 8, o + 6:o + 8] += np.diag(self.ab_reg)
                b[o + 6:o + 8] += self.ab_reg * ab

    def calculate_energy(self):
        le, nv = self.h.evaluate(self.sigma, True, self.fej)
        eps, step = self.h.get_state()
        s = eps + step
        energy = self.energy_marg + self.b_marg @ s + 0.5 * (s @ (self.H_marg @ s))
        for i, m in enumerate(self.meta):
            ab = np.asarray(m["ab0"]) + s[BLOCK * i + 6:BLOCK * i + 8]
            energy += 0.5 * float((ab * self.ab_reg) @ ab)
        return energy + le, nv

    def linearize(self):
        self.H_pose, self.b_pose, self.H_schur, self.b_schur = self.h.linearize(self.sigma, True, self.fej, False)
        eps, _ = self.h.get_state()
        self._prior(self.H_pose, self.b_pose, eps)

    def calculate_step(self, lam):
        eps, _ = self.h.get_state()
        H = self.H_pose + self.H_marg
        b = self.b_pose + self.b_marg
        H[np.diag_indices_from(H)] += np.diag(self.H_pose) * lam
        k = -1.0 / (1.0 + lam)
        H = H + self.H_schur * k
        b = b + self.b_schur * k + self.H_marg @ eps
        step = jacobi_solve(H, b)
        self.h.set_state(None, -step)
        self.h.back_substitute(step, lam)
        return step

    def accept_step(self):
        return self.h.accept()

    def reject_step(self):
        self.h.reject()

    def stop(self):
        return False
