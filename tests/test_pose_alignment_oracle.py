"""Pins oracle/pose_alignment_oracle.py (the coarse tracker's direct image alignment) with property tests, the way the
reference pins EigenPoseAlignment (test/test/energy/problems/test_ceres_pose_alignment.cpp:100-139: the aligner must
reach the ground-truth relative pose) plus a finite-difference check of the 8-parameter Jacobian."""
import numpy as np

from dsopp_b200 import synth
from oracle import pba_oracle as O
from oracle import pose_alignment_oracle as PA


def case_frames(case):
    r, t = case.reference, case.target
    ref = PA.PAFrame(r.T_w_true, r.exposure, r.ab0, r.intr, r.image, r.mask)
    tgt = PA.PAFrame(case.T_w_target_guess, t.exposure, t.ab0, t.intr, t.image, t.mask)
    return ref, tgt


def test_landmarks_from_depth_map_follow_the_reference_constructor():
    case = synth.make_alignment_case(seed=1, width=160, height=120, density=0.2)
    w = case.weight.copy()
    s = case.idepth_sum.copy()
    w[10, 10], s[10, 10] = 2.0, 1e-7  # idepth below kMinIdepth: skipped
    w[2, 50] = 1.0                     # inside the 4-px border: skipped
    uv, idepth, patch = PA.landmarks_from_depth_map(s, w, case.reference.image)
    assert (uv[:, 0] >= 4).all() and (uv[:, 0] <= 160 - 5).all() and (uv[:, 1] >= 4).all() and (uv[:, 1] <= 120 - 5).all()
    assert not ((uv[:, 0] == 10) & (uv[:, 1] == 10)).any()
    order = uv[:, 1] * 160 + uv[:, 0]
    assert (np.diff(order) > 0).all()  # y outer, x inner
    k = 7
    x, y = int(uv[k, 0]), int(uv[k, 1])
    assert idepth[k] == s[y, x] / w[y, x] and patch[k] == case.reference.image[y, x, 0]


def test_jacobian_matches_central_differences_of_the_residual():
    case = synth.make_alignment_case(seed=2, width=160, height=120, density=0.05, ab_scale=1.0)
    # on a linear-ramp image bilinear sampling and the stored gradient are exact (as in the BA property test)
    yy, xx = np.mgrid[0:120, 0:160].astype(np.float64)
    case.reference.image = synth.pixelinfo(20.0 + 0.11 * xx - 0.07 * yy)
    case.target.image = synth.pixelinfo(23.0 + 0.13 * xx - 0.08 * yy)
    ref, tgt = case_frames(case)
    uv, idepth, patch = PA.landmarks_from_depth_map(case.idepth_sum, case.weight, case.reference.image)
    T0 = O.se3_inv(tgt.T_lin) @ ref.T_lin

    def residuals(xi, dab):
        # the update of calculateStep (eigen_pose_alignment.cpp:194-206) with step = -xi on the pose so that J = d r / d xi
        p = PA.PoseAlignerProblem(ref, tgt, uv, idepth, patch, 1e9, (0.0, 0.0), O.se3_exp(xi) @ T0, dab)
        p.calculate_energy()
        s, ab_t = p._scale()
        return (p.t_patch - ab_t[1]) - s * (p.patch - ref.ab0[1]), p.success

    p = PA.PoseAlignerProblem(ref, tgt, uv, idepth, patch, 1e9, (0.0, 0.0), T0)
    p.calculate_energy()
    p.linearize()
    r0, ok0 = residuals(np.zeros(6), np.zeros(2))
    # rebuild the dense Jacobian from H = J^T J columns: compare J^T r and J^T J against finite differences
    J = np.zeros((len(idepth), 8))
    h = 1e-6
    for k in range(8):
        xi, dab = np.zeros(6), np.zeros(2)
        (xi if k < 6 else dab)[k % 6 if k < 6 else k - 6] = h
        rp, okp = residuals(xi, dab)
        rm, okm = residuals(-xi, -dab)
        good = ok0 & okp & okm
        J[good, k] = (rp[good] - rm[good]) / (2 * h)
        J[~good, k] = 0
    good = p.success
    # d_state = -d r / d(left increment) for the pose and +d r / d(ab) for the affine part (signs of :156-169)
    D = np.concatenate([-J[:, :6], J[:, 6:]], axis=1)
    Hfd = D[good].T @ D[good]
    bfd = D[good].T @ r0[good]
    assert np.abs(p.H - Hfd).max() <= 1e-6 * np.abs(Hfd).max()
    assert np.abs(p.b - bfd).max() <= 1e-6 * np.abs(bfd).max()


def test_alignment_recovers_the_ground_truth_relative_pose():
    """test_ceres_pose_alignment.cpp:100-139: distance / angle to the GT relative pose shrink to ~0."""
    # 640x480 sparse (the reference's own regime) and a dense quarter-resolution raster (the configs[2] bound); a
    # 160x120 render of this texture is aliased and has no usable convergence basin for 1-pixel residuals
    for seed, density, W, H in ((3, 0.02, 640, 480), (4, 1.0, 320, 240)):
        case = synth.make_alignment_case(seed=seed, width=W, height=H, density=density, pose_noise=4e-3)
        ref, tgt = case_frames(case)
        uv, idepth, patch = PA.landmarks_from_depth_map(case.idepth_sum, case.weight, case.reference.image)
        trace = []
        out = PA.solve(ref, tgt, uv, idepth, patch, trace=trace)
        T0 = O.se3_inv(tgt.T_lin) @ ref.T_lin
        err0 = np.linalg.norm((O.se3_inv(case.T_t_r_true) @ T0)[:3, 3])
        err1 = np.linalg.norm((O.se3_inv(case.T_t_r_true) @ out["T_t_r"])[:3, 3])
        ang1 = np.arccos(np.clip((np.trace((case.T_t_r_true[:3, :3].T @ out["T_t_r"][:3, :3])) - 1) / 2, -1, 1))
        assert out["n_valid"] > 0.5 * len(idepth)
        assert err1 < 0.1 * err0 and err1 < 3e-4 and ang1 < 1e-4, (seed, err0, err1, ang1)
        e = [t["energy"] for t in trace if t["accepted"]]
        assert all(b < a for a, b in zip(e, e[1:]))  # accepted energies decrease


def test_cpp_restatement_matches_the_numpy_oracle():
    """Two independent restatements of eigen_pose_alignment.cpp (NumPy, vectorised; C++, the reference's serial
    two-pass dataflow) must agree to rounding."""
    for seed, density, W, H, ab_scale, reg in ((3, 0.02, 640, 480, 0.0, (1e12, 1e8)), (6, 0.3, 320, 240, 1.0, (10.0, 1e-2))):
        case = synth.make_alignment_case(seed=seed, width=W, height=H, density=density, pose_noise=4e-3, ab_scale=ab_scale)
        ref, tgt = case_frames(case)
        uv, idepth, patch = PA.landmarks_from_depth_map(case.idepth_sum, case.weight, case.reference.image)
        trace = []
        a = PA.solve(ref, tgt, uv, idepth, patch, ab_reg=reg, trace=trace)
        b = PA.solve_cpp(ref, tgt, uv, idepth, patch, ab_reg=reg)
        assert b["n_valid"] == a["n_valid"] and b["iterations"] == len(trace) and b["converged"] == a["converged"]
        assert abs(b["energy"] - a["energy"]) <= 1e-9 * a["energy"]
        assert np.abs(b["T_t_r"] - a["T_t_r"]).max() <= 1e-9
        assert np.abs(b["ab_eps"] - a["ab_eps"]).max() <= 1e-8 * max(1.0, np.abs(a["ab_eps"]).max())
        assert np.abs(b["H"] - a["H"]).max() <= 1e-9 * np.abs(a["H"]).max()
