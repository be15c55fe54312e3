"""ctypes wrapper of oracle/cpu_ref/pba_cpu_ref.cpp (test infrastructure / timed CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

_libs = {}


def _load(native=False):
    key = "native" if native else "portable"
    if key in _libs:
        return _libs[key]
    path = build_oracle.build_native() if native else build_oracle.build_portable()
    lib = C.CDLL(path)
    lib.cpuref_create.restype = C.c_void_p
    lib.cpuref_create.argtypes = [C.c_int, C.c_int]
    lib.cpuref_destroy.argtypes = [C.c_void_p]
    lib.cpuref_max_threads.restype = C.c_int
    lib.cpuref_push_frame.restype = C.c_int
    lib.cpuref_push_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    lib.cpuref_set_landmarks.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    lib.cpuref_set_statuses.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.cpuref_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cpuref_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cpuref_first_estimate.argtypes = [C.c_void_p]
    lib.cpuref_evaluate.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.cpuref_pose_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.cpuref_schur.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.cpuref_calculate_idepths.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    lib.cpuref_landmarks_energy.restype = C.c_double
    lib.cpuref_landmarks_energy.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.cpuref_accept.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.cpuref_reject.argtypes = [C.c_void_p]
    lib.cpuref_change_statuses.argtypes = [C.c_void_p, C.c_int]
    lib.cpuref_normal_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cpuref_get_residuals.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
    lib.cpuref_get_landmarks.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
    lib.cpuref_gn_iteration.restype = C.c_double
    lib.cpuref_gn_iteration.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_void_p, C.c_double,
                                        C.c_void_p, C.c_void_p]
    lib.cpuref_set_device_ops.argtypes = [C.c_void_p, C.c_int]
    lib.cpuref_update_point_statuses.restype = C.c_double
    lib.cpuref_update_point_statuses.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.cpuref_get_jac_valid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.cpuref_set_idepths.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.cpuref_get_landmark_flags.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _libs[key] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads(native=False):
    return _load(native).cpuref_max_threads()


class CpuWindow:
    def __init__(self, win, use_float=False, threads=1, native=False):
        self.lib = _load(native)
        self.h = C.c_void_p(self.lib.cpuref_create(int(use_float), threads))
        self.n = win.n_frames
        self.counts = [len(f.idepth) for f in win.frames]
        for f in win.frames:
            img = np.ascontiguousarray(f.image, dtype=np.float32)
            mask = np.ascontiguousarray(f.mask, dtype=np.uint8)
            T = np.ascontiguousarray(np.asarray(f.T_w_lin, dtype=np.float64)[:3, :4]).reshape(12)
            ab, it = np.ascontiguousarray(f.ab0, dtype=np.float64), np.ascontiguousarray(f.intr, dtype=np.float64)
            self.lib.cpuref_push_frame(self.h, f.frame_id, _p(img), _p(mask), win.width, win.height, _p(T),
                                       float(f.exposure), _p(ab), _p(it), int(f.fixed))
        for i, f in enumerate(win.frames):
            uv = np.ascontiguousarray(f.uv, dtype=np.float64)
            idp = np.ascontiguousarray(f.idepth, dtype=np.float64)
            pt = np.ascontiguousarray(f.patch, dtype=np.float64)
            fl = np.ascontiguousarray(f.flags, dtype=np.uint8)
            self.lib.cpuref_set_landmarks(self.h, i, len(idp), _p(uv), _p(idp), _p(pt), _p(fl))
        for (r, t), st in win.statuses.items():
            st = np.ascontiguousarray(st, dtype=np.uint8)
            self.lib.cpuref_set_statuses(self.h, r, t, len(st), _p(st))
        eps = np.ascontiguousarray(np.concatenate([f.state_eps for f in win.frames]), dtype=np.float64)
        self.lib.cpuref_set_state(self.h, _p(eps), _p(np.zeros_like(eps)))

        self.frames_meta = [(np.array(f.ab0, dtype=np.float64), bool(f.fixed)) for f in win.frames]

    def close(self):
        if self.h:
            self.lib.cpuref_destroy(self.h)
            self.h = None

    def set_device_ops(self, on=True):
        """float build: residual / energy arithmetic in the CUDA kernels' operation order (bit-exact bookkeeping checks)."""
        self.lib.cpuref_set_device_ops(self.h, int(on))

    def update_point_statuses(self, min_valid, sigma):
        return self.lib.cpuref_update_point_statuses(self.h, int(min_valid), float(sigma))

    def jac_valid(self, r, t):
        out = np.zeros(self.counts[r], np.uint8)
        self.lib.cpuref_get_jac_valid(self.h, r, t, _p(out))
        return out

    def set_statuses(self, r, t, st):
        st = np.ascontiguousarray(st, dtype=np.uint8)
        self.lib.cpuref_set_statuses(self.h, r, t, len(st), _p(st))

    def set_idepths(self, slot, idepth=None, idepth_step=None):
        a = None if idepth is None else np.ascontiguousarray(idepth, dtype=np.float64)
        b = None if idepth_step is None else np.ascontiguousarray(idepth_step, dtype=np.float64)
        self.lib.cpuref_set_idepths(self.h, slot, self.counts[slot], _p(a), _p(b))

    def landmark_flags(self, slot):
        n = self.counts[slot]
        out = dict(outlier=np.zeros(n, np.uint8), n_inliers=np.zeros(n, np.uint32), rel_baseline=np.zeros(n))
        self.lib.cpuref_get_landmark_flags(self.h, slot, _p(out["outlier"]), _p(out["n_inliers"]), _p(out["rel_baseline"]))
        return out

    def __del__(self):
        self.close()

    def first_estimate(self):
        self.lib.cpuref_first_estimate(self.h)

    def evaluate(self, sigma, fej, eval_jac, huber=True):
        self.lib.cpuref_evaluate(self.h, sigma, int(fej), int(eval_jac), int(huber))

    def pose_pose(self, for_marg=False):
        d = 8 * self.n
        H, b = np.zeros((d, d)), np.zeros(d)
        self.lib.cpuref_pose_pose(self.h, int(for_marg), _p(H), _p(b))
        return H, b

    def schur(self, for_marg=False):
        d = 8 * self.n
        H, b = np.zeros((d, d)), np.zeros(d)
        self.lib.cpuref_schur(self.h, int(for_marg), _p(H), _p(b))
        return H, b

    def calculate_idepths(self, step, lam):
        s = np.ascontiguousarray(step, dtype=np.float64)
        self.lib.cpuref_calculate_idepths(self.h, _p(s), lam)

    def landmarks_energy(self, for_marg=False):
        n = C.c_int()
        e = self.lib.cpuref_landmarks_energy(self.h, int(for_marg), C.byref(n))
        return e, n.value

    def set_state(self, eps=None, step=None):
        eps = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        step = None if step is None else np.ascontiguousarray(step, dtype=np.float64)
        self.lib.cpuref_set_state(self.h, _p(eps), _p(step))

    def get_state(self):
        eps, step = np.zeros(8 * self.n), np.zeros(8 * self.n)
        self.lib.cpuref_get_state(self.h, _p(eps), _p(step))
        return eps, step

    def accept(self):
        a, b = C.c_double(), C.c_double()
        self.lib.cpuref_accept(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def reject(self):
        self.lib.cpuref_reject(self.h)

    def change_statuses(self, accept=True):
        self.lib.cpuref_change_statuses(self.h, int(accept))

    def residuals(self, r, t):
        n = self.counts[r]
        out = dict(r=np.zeros((n, 8)), J_ref=np.zeros((n, 8, 8)), J_tgt=np.zeros((n, 8, 8)), d_idepth=np.zeros((n, 8)),
                   w=np.zeros(n), e=np.zeros(n), status=np.zeros(n, np.uint8), cand=np.zeros(n, np.uint8))
        self.lib.cpuref_get_residuals(self.h, r, t, _p(out["r"]), _p(out["J_ref"]), _p(out["J_tgt"]),
                                      _p(out["d_idepth"]), _p(out["w"]), _p(out["e"]), _p(out["status"]),
                                      _p(out["cand"]))
        return out

    def landmarks(self, slot):
        n, d = self.counts[slot], 8 * self.n
        out = dict(idepth=np.zeros(n), idepth_step=np.zeros(n), inv_hdd=np.zeros(n), b_d=np.zeros(n),
                   ill=np.zeros(n, np.uint8), hpd=np.zeros((n, d)))
        self.lib.cpuref_get_landmarks(self.h, slot, _p(out["idepth"]), _p(out["idepth_step"]), _p(out["inv_hdd"]),
                                      _p(out["b_d"]), _p(out["ill"]), _p(out["hpd"]), d)
        return out

    def gn_iteration(self, sigma, fej, lam, ab_reg, fixed_reg):
        """One LM loop body (linearize, step, energy, accept); returns (energy, times[6], step)."""
        times = np.zeros(6)
        step = np.zeros(8 * self.n)
        reg = np.ascontiguousarray(ab_reg, dtype=np.float64)
        e = self.lib.cpuref_gn_iteration(self.h, sigma, int(fej), lam, _p(reg), fixed_reg, _p(times), _p(step))
        return e, times, step


class CpuRefProblem:
    """LevenbergMarquardtProblem over a CpuWindow (PBA/eigen_photometric_bundle_adjustment_problem.hpp:255-429): the same
    six methods as oracle.pba_oracle.Problem, so oracle.pba_oracle.lm_solve drives it -- the C++ restatement at sizes the
    NumPy oracle would need minutes for.  Priors and the reduced solve are done here in float64 NumPy."""

    def __init__(self, cw: CpuWindow, sigma, ab_reg=(1e12, 1e8), fixed_reg=1e16, fej=True, H_marg=None, b_marg=None,
                 energy_marg=0.0):
        self.cw, self.sigma, self.fej = cw, float(sigma), fej
        self.ab_reg, self.fixed_reg = np.asarray(ab_reg, dtype=np.float64), float(fixed_reg)
        d = 8 * cw.n
        self.H_marg = np.zeros((d, d)) if H_marg is None else np.array(H_marg, dtype=np.float64)
        self.b_marg = np.zeros(d) if b_marg is None else np.array(b_marg, dtype=np.float64)
        self.energy_marg = float(energy_marg)

    def calculate_energy(self):
        self.cw.evaluate(self.sigma, self.fej, False, True)
        eps, step = self.cw.get_state()
        s = eps + step
        energy = self.energy_marg + self.b_marg @ s + 0.5 * (s @ (self.H_marg @ s))
        for i, (ab0, _) in enumerate(self.cw.frames_meta):
            ab = ab0 + s[8 * i + 6:8 * i + 8]
            energy += 0.5 * float((ab * self.ab_reg) @ ab)
        le, nv = self.cw.landmarks_energy()
        return energy + le, nv

    def linearize(self):
        self.cw.evaluate(self.sigma, self.fej, True, True)
        self.H_pose, self.b_pose = self.cw.pose_pose()
        eps, _ = self.cw.get_state()
        for i, (ab0, fixed) in enumerate(self.cw.frames_meta):
            o = 8 * i
            if fixed:
                self.H_pose[o:o + 8, o:o + 8] += np.eye(8) * self.fixed_reg
                self.b_pose[o:o + 8] += self.fixed_reg * eps[o:o + 8]
            else:
                self.H_pose[o + 6:o + 8, o + 6:o + 8] += np.diag(self.ab_reg)
                self.b_pose[o + 6:o + 8] += self.ab_reg * (ab0 + eps[o + 6:o + 8])
        self.H_schur, self.b_schur = self.cw.schur()

    def calculate_step(self, lam):
        eps, _ = self.cw.get_state()
        H = self.H_pose + self.H_marg
        b = self.b_pose + self.b_marg
        H[np.diag_indices_from(H)] += np.diag(self.H_pose) * lam
        k = -1.0 / (1.0 + lam)
        H = H + self.H_schur * k
        b = b + self.b_schur * k + self.H_marg @ eps
        step = normal_solve(H, b)
        self.cw.set_state(None, -step)
        self.cw.calculate_idepths(step, lam)
        return step

    def accept_step(self):
        return self.cw.accept()

    def reject_step(self):
        self.cw.reject()

    def stop(self):
        return False


def normal_solve(H, b, native=False):
    lib = _load(native)
    H = np.ascontiguousarray(H, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    lib.cpuref_normal_solve(len(b), _p(H), _p(b), _p(x))
    return x
