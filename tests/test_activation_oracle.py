"""Property tests of oracle/activation_oracle.py (landmarks_activator.cpp:122-316): the 1-D refine pulls a perturbed
inverse depth towards the true one, the derivative is the derivative, and the delete rules fire."""
import numpy as np

from dsopp_b200 import synth
from oracle import activation_oracle as A
from oracle import pba_oracle as O


def frames_of(win):
    return [A.ActFrame(f.frame_id, f.T_w_true, f.exposure, f.ab0, f.intr, f.image, f.mask) for f in win.frames]


def test_refine_improves_perturbed_idepths_and_deletes_hopeless_ones():
    win = synth.make_window(n_frames=5, points_per_frame=60, seed=21, pose_noise=0.0, eps_scale=0.0, ab_scale=0.0)
    fr = frames_of(win)
    rng = np.random.default_rng(0)
    better, total = 0, 0
    for r in (0, 2):
        f = win.frames[r]
        for l in range(60):
            rho0 = f.idepth_true[l] * (1 + rng.uniform(-1, 1) * 0.03)
            act, rho, n = A.optimize_immature_landmark(fr[r], fr, f.uv[l], f.patch[l], rho0, 3, 20.0)
            if act:
                total += 1
                better += abs(rho - f.idepth_true[l]) < abs(rho0 - f.idepth_true[l])
                assert n >= 3 and rho > 0
    assert total > 100 and better > 0.9 * total
    # a landmark that reprojects outside every target is deleted, with idepth = -1 (stop_ path, :153-156,192-195)
    act, rho, n = A.optimize_immature_landmark(fr[0], fr, np.array([20.0, 20.0]), win.frames[0].patch[0], 5.0, 1, 20.0)
    assert not act and (rho == -1.0 or n == 0)


def test_hessian_and_b_are_the_derivatives_of_the_energy_terms():
    win = synth.make_window(n_frames=3, points_per_frame=10, seed=22, pose_noise=0.0, eps_scale=0.0, ab_scale=0.0)
    H, W = win.height, win.width
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    for k, f in enumerate(win.frames):
        f.image = synth.pixelinfo(20.0 + 0.11 * xx - 0.07 * yy + 0.3 * k)  # exact bilinear gradients
        f.exposure = 1.0
    fr = frames_of(win)
    f = win.frames[0]
    rho = f.idepth_true[3]
    pat = (f.uv[3][None, :] + O.PATTERN).astype(int)
    f.patch[3] = fr[0].image[pat[:, 1], pat[:, 0], 0] + 1.5  # consistent with the ramp, small non-zero residuals
    p = A.LandmarkActivationProblem(fr[0], fr, f.uv[3], f.patch[3], 1e9, rho)  # sigma huge: plain least squares
    p.linearize()

    def half_sq(r_):
        q = A.LandmarkActivationProblem(fr[0], fr, f.uv[3], f.patch[3], 1e9, r_)
        e, n = q.calculate_energy()
        return 0.5 * e, n
    h = 1e-6
    (ep, n1), (em, n2), (e0, n0) = half_sq(rho + h), half_sq(rho - h), half_sq(rho)
    assert n0 == n1 == n2 == 2
    assert abs((ep - em) / (2 * h) - p.b) <= 1e-5 * max(1.0, abs(p.b))            # d(1/2 sum r^2)/d rho = sum d r
    assert abs((ep - 2 * e0 + em) / (h * h) - p.hessian) <= 2e-2 * p.hessian       # Gauss-Newton Hessian (ramp image: exact up to the projection's curvature)
