"""GPU parity of the coarse-tracker direct image alignment: the one-cluster CUDA LM solve, called
through the C ABI (include/dsopp_cuda_pose_alignment.h), against the float64 NumPy oracle on the same inputs.

Tolerances (fp32 per-point arithmetic vs float64): energy 2e-4 relative, same number of valid residuals up to the
points whose reprojection lies within fp32 rounding of the ROI border, final t_t_r within 2e-5 (translation, metres at
depth ~5) / 2e-5 rad, iteration counts within 1 (the accept / convergence tests compare energies that agree to ~1e-5)."""
import numpy as np
import pytest

from dsopp_b200 import synth

pytestmark = pytest.mark.gpu


def run_case(seed, W, H, density, ab_scale=0.0, device_depth_map=False, mask_hole=False, ab_reg=(1e12, 1e8),
             grid_threshold=None):
    from dsopp_b200 import pose_alignment as G
    from oracle import pba_oracle as O
    from oracle import pose_alignment_oracle as PA
    case = synth.make_alignment_case(seed=seed, width=W, height=H, density=density, pose_noise=4e-3, ab_scale=ab_scale)
    r, t = case.reference, case.target
    mask = t.mask.copy()
    if mask_hole:
        mask[H // 3:H // 2, W // 4:W // 2] = 0
    ref = PA.PAFrame(r.T_w_true, r.exposure, r.ab0, r.intr, r.image, r.mask)
    tgt = PA.PAFrame(case.T_w_target_guess, t.exposure, t.ab0, t.intr, t.image, mask)
    ids32, w32 = case.idepth_sum.astype(np.float32), case.weight.astype(np.float32)
    uv, idepth, patch = PA.landmarks_from_depth_map(ids32.astype(np.float64), w32.astype(np.float64), r.image)
    trace = []
    want = PA.solve(ref, tgt, uv, idepth, patch, ab_reg=ab_reg, trace=trace)
    al = G.Aligner(max(len(idepth), 1), W, H)
    if device_depth_map:
        n = al.set_reference_depth_map(r.image, ids32, w32, r.T_w_true, r.exposure, r.ab0, r.intr)
        assert n == len(idepth)
        xy_d, id_d, pt_d = al.get_reference_landmarks()
        assert (xy_d == uv.astype(np.float32)).all()  # same landmarks, same order (y outer, x inner): bit-exact bookkeeping
        assert np.allclose(id_d, idepth, rtol=1e-6) and (pt_d == patch.astype(np.float32)).all()
    else:
        al.set_reference_landmarks(uv, idepth, patch, r.T_w_true, r.exposure, r.ab0, r.intr, W, H)
    al.set_target(t.image, mask, case.T_w_target_guess, t.exposure, t.ab0, t.intr)
    if grid_threshold is not None:
        al.set_grid_threshold(grid_threshold)
    got = al.solve(G.default_options(ab_reg=ab_reg))
    gtr = al.trace()
    al.close()
    print("oracle trace:", [(round(t_["energy"], 4), t_["accepted"]) for t_ in trace])
    print("gpu trace:   ", [(round(t_["energy"], 4), t_["accepted"]) for t_ in gtr])
    print(f"[pa seed {seed} {W}x{H} n={len(idepth)}] oracle E={want['energy']:.4f} n={want['n_valid']} it={len(trace)} | "
          f"gpu E={got['energy']:.4f} n={got['n_valid']} it={got['iterations']} rmse {got['rmse']:.4f}")
    assert abs(got["n_valid"] - want["n_valid"]) <= max(2, int(2e-4 * want["n_valid"]))
    assert abs(got["energy"] - want["energy"]) <= 2e-4 * want["energy"]
    assert abs(got["rmse"] - want["rmse"]) <= 2e-4 * want["rmse"]
    # The traces must agree while the energy still moves: every iteration in which the oracle's trial energy differs
    # from its current energy by more than 1e-4 (relative) has the same decision and the same energy on the GPU.  Near
    # the minimum the trial energies sit within fp32 noise (~1e-5 relative) of each other, where the function-tolerance
    # test (1e-5) and "E1 < E0" are legitimately ambiguous, so the iteration COUNT may differ there.
    cur = None
    for k, (a, b) in enumerate(zip(gtr, trace)):
        base = cur if cur is not None else b["energy"] * 2
        if abs(b["energy"] - base) <= 1e-4 * base:
            break
        assert a["accepted"] == b["accepted"], k
        assert abs(a["energy"] - b["energy"]) <= 2e-4 * b["energy"], k
        cur = b["energy"] if b["accepted"] else cur
    assert got["iterations"] <= 50
    d = O.se3_inv(want["T_t_r"]) @ got["T_t_r"]
    ang = np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))
    assert np.linalg.norm(d[:3, 3]) <= 2e-5 and ang <= 2e-5, (np.linalg.norm(d[:3, 3]), ang)
    assert np.abs(got["ab_eps"] - want["ab_eps"]).max() <= 1e-4 * max(1.0, np.abs(want["ab_eps"]).max()) + 1e-6
    assert np.abs(got["T_w_target"] - want["T_w_target"]).max() <= 5e-5
    Hs = np.abs(want["H"]).max()
    assert np.abs(got["H"] - want["H"]).max() <= 2e-4 * Hs
    # and the point of the exercise: the ground-truth relative pose is recovered
    e = O.se3_inv(case.T_t_r_true) @ got["T_t_r"]
    assert np.linalg.norm(e[:3, 3]) < 3e-4
    return got


def test_sparse_depth_map_full_resolution():
    run_case(seed=3, W=640, H=480, density=0.02)


def test_dense_raster_quarter_resolution_with_device_side_landmarks():
    run_case(seed=4, W=320, H=240, density=1.0, device_depth_map=True)


def test_dense_raster_full_resolution_on_the_whole_chip():
    """BASELINE configs[2] as stated: 640x480, every pixel carries depth (298 k one-pixel residuals), 6 + 2 parameters.
    298 k landmarks select the cooperative whole-chip kernel (one CTA per SM, grid barrier per sweep)."""
    run_case(seed=5, W=640, H=480, density=1.0, device_depth_map=True)


def test_whole_chip_and_cluster_kernels_agree():
    """Same arithmetic per point; what differs is which points share a thread's fp32 running sums (the stride is the
    number of CTAs x 512) and the order of the CTAs' fp64 partials: agreement to fp32 summation noise (1e-6 relative)."""
    a = run_case(seed=4, W=320, H=240, density=1.0, device_depth_map=True, grid_threshold=0)    # whole chip
    b = run_case(seed=4, W=320, H=240, density=1.0, device_depth_map=True, grid_threshold=-1)   # one cluster
    assert a["iterations"] == b["iterations"] and a["n_valid"] == b["n_valid"]
    assert abs(a["energy"] - b["energy"]) <= 2e-6 * abs(b["energy"])
    assert np.abs(a["T_t_r"] - b["T_t_r"]).max() <= 1e-6
    assert np.abs(a["H"] - b["H"]).max() <= 2e-6 * np.abs(b["H"]).max()


def test_affine_brightness_is_estimated_when_the_prior_is_weak():
    got = run_case(seed=6, W=320, H=240, density=0.3, ab_scale=1.0, ab_reg=(10.0, 1e-2), device_depth_map=True)
    assert np.abs(got["ab_eps"]).max() > 0


def test_target_mask_with_a_hole():
    run_case(seed=7, W=320, H=240, density=0.3, mask_hole=True)


def test_empty_reference_and_error_codes():
    from dsopp_b200 import capi, pose_alignment as G
    case = synth.make_alignment_case(seed=8, width=160, height=120, density=0.1)
    r, t = case.reference, case.target
    al = G.Aligner(100, 160, 120)
    with pytest.raises(capi.DpbaError):
        al.solve()  # frames not pushed
    al.set_reference_landmarks(np.zeros((0, 2)), np.zeros(0), np.zeros(0), r.T_w_true, r.exposure, r.ab0, r.intr, 160, 120)
    al.set_target(t.image, t.mask, case.T_w_target_guess, t.exposure, t.ab0, t.intr)
    out = al.solve()
    assert out["n_valid"] == 0 and out["iterations"] == 0  # the LM loop is not entered without residuals
    with pytest.raises(capi.DpbaError):
        al.set_reference_landmarks(np.zeros((101, 2)), np.zeros(101), np.zeros(101), r.T_w_true, r.exposure, r.ab0, r.intr, 160, 120)
    al.close()


def test_coarse_to_fine_through_the_cpp_host_class():
    """The tracker's loop (monocular_tracker.cpp:199-214) over 4 pyramid levels, coarse to fine, through the C++
    CudaPoseAlignment mirror of EigenPoseAlignment: reset / pushFrame x2 / solve per level, each level starting from
    the previous level's pose.  Level images are 2x2 box pyramids (downscale_image.hpp:16-33), per-level intrinsics
    are f / 2^l without a half-pixel shift (camera_calibration.cpp:66-70), depth-map accumulators are summed 2x2
    (create_depth_maps.cpp)."""
    from dsopp_b200 import host
    from oracle import pba_oracle as O
    case = synth.make_alignment_case(seed=9, width=640, height=480, density=0.05, pose_noise=1.5e-2)
    r, t = case.reference, case.target
    levels = 4
    ref_I, tgt_I = [r.image[..., 0]], [t.image[..., 0]]
    ids, w = [case.idepth_sum.astype(np.float32)], [case.weight.astype(np.float32)]
    for _ in range(1, levels):
        ref_I.append(synth.downscale(ref_I[-1]))
        tgt_I.append(synth.downscale(tgt_I[-1]))
        a, b = ids[-1], w[-1]
        ids.append(a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])
        w.append(b[0::2, 0::2] + b[1::2, 0::2] + b[0::2, 1::2] + b[1::2, 1::2])
    al = host.PoseAligner(640, 480)
    T = case.T_w_target_guess
    err0 = np.linalg.norm((O.se3_inv(case.T_t_r_true) @ (O.se3_inv(T) @ r.T_w_true))[:3, 3])
    last = None
    for lvl in range(levels - 1, -1, -1):
        intr = r.intr / (2 ** lvl)
        Hh, Ww = ref_I[lvl].shape
        out = al.align_level(intr, synth.pixelinfo(ref_I[lvl]), ids[lvl], w[lvl], r.T_w_true, r.exposure, r.ab0,
                             synth.pixelinfo(tgt_I[lvl]), np.full((Hh, Ww), 255, np.uint8), T, t.exposure, t.ab0)
        assert out["n"] > 0 and out["rmse"] > 0
        T = out["T_w_target"]
        last = out
        err = np.linalg.norm((O.se3_inv(case.T_t_r_true) @ (O.se3_inv(T) @ r.T_w_true))[:3, 3])
        print(f"[coarse-to-fine] level {lvl}: {Ww}x{Hh}, {out['n']} landmarks, rmse {out['rmse']:.3f}, |dt| {err:.2e}")
    al.close()
    assert err < 0.05 * err0 and err < 5e-4
    cov = last["cov"]
    assert np.allclose(cov, cov.T, atol=1e-12 * np.abs(cov).max()) and (np.diag(cov) > 0).all()
