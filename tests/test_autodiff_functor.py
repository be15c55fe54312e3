"""Second, independent pin of the oracle: the reference's AUTODIFF cost functor restated from scratch.

The reference cross-checks its analytic Jacobians (PBA/evaluate_jacobians.hpp) against Ceres autodiff of
BundleAdjustmentPhotometricCostFunctor
(src/energy/problems/internal/energy/problems/cost_functors/bundle_adjustment_photometric_cost_functor.hpp:27-172)
in test/test/energy/problems/test_analytical_diff.cpp:49-156 (each norm < 1e-5, zero state eps, FEJ on, two
frames).  Ceres is absent here, so the functor is written again below in torch float64 -- following ONLY that
file, not oracle/pba_oracle.py -- and differentiated by autograd:

  * t_target_reference = t_t_r0.rightIncrement(eps_r).leftIncrement(-eps_t)            (:89)
  * affine brightness = brightness0 + eps                                              (:90-91)
  * reprojectPattern through pinhole models built from the intrinsics                  (:95-101)
  * Grid2D::Evaluate on Jets: value = bilinear I, derivative = bilinear STORED (dx, dy)
    (src/features/include/features/camera/pixel_map.hpp:245-258)
  * residual = (I_t - b_t) - (tau_t / tau_r) exp(a_t - a_r) (patch - b_r)              (:110-118)

As the reference's comment says (test_analytical_diff.cpp:108-113) the two derivatives agree only at zero eps.
"""
import numpy as np
import pytest
import torch

from dsopp_b200 import synth
from oracle import pba_oracle as O

DT = torch.float64


def hat(w):
    z = torch.zeros((), dtype=DT)
    return torch.stack([torch.stack([z, -w[2], w[1]]), torch.stack([w[2], z, -w[0]]), torch.stack([-w[1], w[0], z])])


def se3_exp(xi):
    """Sophus::SE3::exp, tangent = [upsilon; omega] (Sophus @593db475, published closed form); series at theta -> 0
    so that autograd sees the exact derivative at xi = 0."""
    v, w = xi[:3], xi[3:]
    th2 = w @ w
    W = hat(w)
    W2 = W @ W
    # a = sin(th)/th, b = (1-cos th)/th^2, c = (th - sin th)/th^3 as even power series in th (exact to 1e-16 for th < 1e-2)
    a = 1 - th2 / 6 + th2 * th2 / 120
    b = 0.5 - th2 / 24 + th2 * th2 / 720
    c = 1.0 / 6 - th2 / 120 + th2 * th2 / 5040
    eye = torch.eye(3, dtype=DT)
    R = eye + a * W + b * W2
    V = eye + b * W + c * W2
    T = torch.eye(4, dtype=DT)
    T = T.clone()
    T[:3, :3] = R
    T[:3, 3] = V @ v
    return T


class Grid(torch.autograd.Function):
    """PixelMap<1>::Evaluate(Jet r, Jet c, Jet* f): f.a = interpolated I, f.v = dx * c.v + dy * r.v."""

    @staticmethod
    def forward(ctx, u, v, image):
        img = image.numpy()
        uu, vv = u.detach().numpy(), v.detach().numpy()
        ix, iy = uu.astype(np.int64), vv.astype(np.int64)  # truncation, pixel_map.hpp:22-23
        dx, dy = uu - ix, vv - iy
        w11, w10, w01 = dx * dy, dy - dx * dy, dx - dx * dy
        w00 = 1 - dx - dy + dx * dy
        val = (w11[:, None] * img[iy + 1, ix + 1] + w10[:, None] * img[iy + 1, ix] + w01[:, None] * img[iy, ix + 1]
               + w00[:, None] * img[iy, ix])
        ctx.save_for_backward(torch.from_numpy(val[:, 1].copy()), torch.from_numpy(val[:, 2].copy()))
        return torch.from_numpy(val[:, 0].copy())

    @staticmethod
    def backward(ctx, g):
        gx, gy = ctx.saved_tensors
        return g * gx, g * gy, None


def functor(ref, tgt, l, eps_r, eps_t, idepth):
    """BundleAdjustmentPhotometricCostFunctor::operator() for landmark l of `ref` seen in `tgt` (zero-based eps)."""
    T0 = torch.from_numpy(np.linalg.inv(tgt.T_lin) @ ref.T_lin)                    # t_t_r0, :58
    T = se3_exp(-eps_t[:6]) @ T0 @ se3_exp(eps_r[:6])                              # :89
    ab_r = torch.from_numpy(ref.ab0) + eps_r[6:]
    ab_t = torch.from_numpy(tgt.ab0) + eps_t[6:]
    fx, fy, cx, cy = [float(x) for x in ref.intr]
    pat = torch.from_numpy(ref.uv[l][None, :] + O.PATTERN)                          # reference_pattern
    ray = torch.stack([(pat[:, 0] - cx) / fx, (pat[:, 1] - cy) / fy, torch.ones(8, dtype=DT)], dim=1)
    p = ray @ T[:3, :3].T + idepth * T[:3, 3]                                       # R K^-1 [u v 1] + rho t
    fx, fy, cx, cy = [float(x) for x in tgt.intr]
    u = fx * p[:, 0] / p[:, 2] + cx
    v = fy * p[:, 1] / p[:, 2] + cy
    I = Grid.apply(u, v, torch.from_numpy(np.ascontiguousarray(tgt.image)))
    scale = (tgt.exposure / ref.exposure) * torch.exp(ab_t[0] - ab_r[0])
    return (I - ab_t[1]) - scale * (torch.from_numpy(ref.patch[l]) - ab_r[1])       # SimilarityMeasureSSD: a - b


def autodiff(ref, tgt, l):
    z = torch.zeros(8, dtype=DT)
    rho = torch.tensor(float(ref.idepth[l]), dtype=DT)
    r = functor(ref, tgt, l, z, z, rho)
    J = torch.autograd.functional.jacobian(lambda a, b, c: functor(ref, tgt, l, a, b, c), (z, z, rho))
    return r.numpy(), J[0].numpy(), J[1].numpy(), J[2].numpy()


def oracle_frames(n_frames, fej, seed):
    win = synth.make_window(n_frames=n_frames, points_per_frame=12, seed=seed, eps_scale=0.0, ab_scale=1.0)
    frames = O.frames_from_window(win)
    for f in frames:
        assert np.all(f.state_eps == 0)
    O.first_estimate_jacobians(frames)
    O.evaluate_jacobians(frames, 0.0, fej=fej, evaluate_jacobians=True, new_point=True, huber=False)
    return frames


@pytest.mark.parametrize("n_frames,fej", [(2, True), (2, False), (4, False)])
def test_analytic_jacobians_equal_autodiff_of_the_reference_functor(n_frames, fej):
    """test_analytical_diff.cpp:151-155 -- residuals, d_idepth, d_host_state_eps, d_target_state_eps, norms < 1e-5.
    FEJ on only with two frames (as in the reference test): with more, quirk Q1 makes the `a` column of every target
    use the brightness scale towards the LAST target, which autodiff of a single pair cannot (and should not) show."""
    frames = oracle_frames(n_frames, fej, seed=11)
    checked = 0
    for ref in frames:
        for tgt in frames:
            if ref is tgt:
                continue
            res = ref.residuals[tgt.id]
            for l in range(len(ref.idepth)):
                if res.cand[l] != O.K_OK:
                    continue
                r, Jr, Jt, Jd = autodiff(ref, tgt, l)
                assert np.linalg.norm(res.r[l] - r) < 1e-5 * max(1.0, np.linalg.norm(r))
                scale = max(1.0, np.abs(Jr).max())  # pose columns are O(1e3): the reference bound, made relative
                assert np.linalg.norm(res.J_ref[l] - Jr) < 1e-5 * scale
                assert np.linalg.norm(res.J_tgt[l] - Jt) < 1e-5 * scale
                assert np.linalg.norm(res.d_idepth[l] - Jd) < 1e-5 * max(1.0, np.abs(Jd).max())
                checked += 1
    assert checked >= 12 * (n_frames - 1)


def test_q1_last_target_scale_is_what_differs_under_fej_with_three_frames():
    """Quirk Q1 made explicit: with FEJ and >2 frames only the affine `a` columns deviate from autodiff, and by exactly
    the ratio of brightness scales (first_estimate_jacobians.hpp:23,57-63)."""
    frames = oracle_frames(3, True, seed=12)
    ref, tgt, last = frames[0], frames[1], frames[2]
    res = ref.residuals[tgt.id]
    l = int(np.flatnonzero(res.cand == O.K_OK)[0])
    _, Jr, Jt, _ = autodiff(ref, tgt, l)
    assert np.allclose(res.J_ref[l][:, :6], Jr[:, :6], rtol=0, atol=1e-5 * np.abs(Jr).max())
    assert np.allclose(res.J_ref[l][:, 7], Jr[:, 7], rtol=1e-9)
    s_pair = (tgt.exposure / ref.exposure) * np.exp(tgt.ab0[0] - ref.ab0[0])
    s_last = (last.exposure / ref.exposure) * np.exp(last.ab0[0] - ref.ab0[0])
    assert np.allclose(res.J_ref[l][:, 6], Jr[:, 6] * s_last / s_pair, rtol=1e-9)
    assert np.allclose(res.J_tgt[l][:, 6], Jt[:, 6] * s_last / s_pair, rtol=1e-9)
