"""GPU checks of the device-side paths next to the solve: the exact radix-select quantile of updatePointStatuses against the
host nth_element path, the reference depth maps of the coarse tracker and the mean-square optical flow of the keyframe
decision against their oracles.  (Round 1 wrote them after its GPU minutes were spent; they first ran -- and passed -- on a
B200 in round 2, which is when the device quantile became the default.)
"""
import numpy as np
import pytest

from dsopp_b200 import synth

pytestmark = pytest.mark.gpu

SIGMA = 20.0


@pytest.mark.parametrize("marginalize_first", [False, True])
def test_device_quantile_equals_host_nth_element(marginalize_first):
    """updatePointStatuses (photometric_bundle_adjustment.cpp:325-405): the radix select on the device must return the
    very float std::nth_element returns on the host, and leave identical statuses / flags / inlier counts."""
    from dsopp_b200 import capi

    win = synth.make_window(n_frames=5, points_per_frame=300, seed=5, marginalize_first=marginalize_first)
    out = []
    for dev in (0, 1):
        h = capi.upload_window(win)
        h.set_option("device_quantile", dev)
        h.first_estimate()
        h.evaluate(SIGMA, True, True)
        h.change_residual_statuses(True)
        thr = h.update_point_statuses(1, SIGMA)
        lms = [h.get_landmarks(i) for i in range(win.n_frames)]
        sts = [h.get_frame_statuses(i) for i in range(win.n_frames)]
        out.append((thr, lms, sts))
        h.close()
    (thr0, lm0, st0), (thr1, lm1, st1) = out
    assert thr0 == thr1
    for a, b in zip(lm0, lm1):
        for k in ("flags", "n_inliers", "rel_baseline"):
            assert np.array_equal(a[k], b[k]), k
    for a, b in zip(st0, st1):
        assert np.array_equal(np.asarray(a), np.asarray(b))


def test_device_quantile_with_no_eligible_residual():
    from dsopp_b200 import capi

    win = synth.make_window(n_frames=3, points_per_frame=50, seed=6)
    for f in win.frames:
        f.flags[:] = synth.FLAG_MARGINALIZED
    h = capi.upload_window(win)
    h.set_option("device_quantile", 1)
    h.first_estimate()
    h.evaluate(SIGMA, True, True)
    h.change_residual_statuses(True)
    assert h.update_point_statuses(1, SIGMA) == 0.0
    h.close()


def _compare_maps(got, ref, rtol):
    for lvl, ((gi, gw), (ri, rw)) in enumerate(zip(got, ref)):
        assert gi.shape == ri.shape
        g_on, r_on = gw > 0, rw > 0
        # fp32 reprojection on the device vs float64 in the oracle: a landmark whose reprojection lies within ~1e-4 px of
        # a rounding boundary may land on the neighbouring pixel (and drag its dilated neighbours along)
        assert (g_on != r_on).sum() <= 12 * (1 + (lvl == 0)), (lvl, (g_on != r_on).sum())
        both = g_on & r_on
        bad_w = np.abs(gw[both] - rw[both]) > rtol * np.abs(rw[both])
        bad_i = np.abs(gi[both] - ri[both]) > rtol * np.abs(ri[both])
        assert bad_w.sum() <= 12 and bad_i.sum() <= 12, (lvl, bad_w.sum(), bad_i.sum())


def test_reference_depth_maps_equal_the_oracle():
    """createReferenceDepthMaps (create_depth_maps.cpp:122-146) on the device vs oracle/depth_map_oracle.py."""
    from dsopp_b200 import capi
    from oracle import depth_map_oracle as D
    from oracle import pba_oracle as O

    win = synth.make_window(n_frames=5, points_per_frame=400, seed=8, ab_scale=0.0)
    win.frames[1].flags[::7] |= synth.FLAG_OUTLIER
    win.statuses[(2, 4)][::5] = 1  # kOutlier towards the newest keyframe
    frames = O.frames_from_window(win)
    h = capi.upload_window(win)
    h.first_estimate()
    # constant variance (the reference without estimate_uncertainty)
    got = h.create_reference_depth_maps(4, 1e-5)
    ref = D.create_reference_depth_maps(frames, 4)
    _compare_maps(got, ref, 1e-4)
    # per-landmark variance = inv_hessian_idepth_idepth of the last linearisation
    O.first_estimate_jacobians(frames)
    prob = O.Problem(frames, SIGMA)
    prob.linearize()
    h.linearize(SIGMA, True, True, False)
    got = h.create_reference_depth_maps(4, -1.0)
    ref = D.create_reference_depth_maps(frames, 4, [f.inv_hdd for f in frames])
    _compare_maps(got, ref, 5e-3)
    h.close()


def test_mean_square_optical_flow_equals_the_oracle():
    """calculateMeanSquareOpticalFlow (monocular_tracker.cpp:104-133) over the aligner's resident landmark list."""
    from dsopp_b200 import pose_alignment
    from oracle import depth_map_oracle as D
    from oracle import pba_oracle as O
    from oracle import pose_alignment_oracle as PA

    win = synth.make_window(n_frames=5, points_per_frame=500, seed=11, pose_noise=0.0, idepth_noise=0.0, eps_scale=0.0,
                            ab_scale=0.0)
    frames = O.frames_from_window(win)
    idw, wgt = D.create_reference_depth_maps(frames, 1)[0]
    sf = win.frames[-1]
    al = pose_alignment.Aligner(idw.size, win.width, win.height)
    n = al.set_reference_depth_map(sf.image, idw.astype(np.float32), wgt.astype(np.float32), sf.T_w_lin, sf.exposure, sf.ab0,
                                   sf.intr)
    assert n > 5000
    for xi in ([0.03, -0.02, 0.05, 0.01, -0.015, 0.008], [0.03, -0.02, 0.05, 0, 0, 0], [1.5, 0.4, 0.0, 0.0, 0.3, 0.0]):
        T = O.se3_exp(np.array(xi))
        ref_flow, ref_n = PA.mean_square_optical_flow(idw, wgt, T, np.asarray(sf.intr, np.float64))
        flow, used = al.mean_square_optical_flow(T)
        assert abs(used - ref_n) <= 3 and abs(flow - ref_flow) <= 2e-4 * ref_flow
    flow, used = al.mean_square_optical_flow(O.se3_exp(np.array([0, 0, -30.0, 0, 0, 0])))
    assert used == 0 and np.isnan(flow)
    al.close()


@pytest.mark.parametrize("shape", [(640, 480), (320, 240), (88, 60)])
def test_tma_staged_pixelinfo_is_bit_identical_to_the_direct_stencil(shape):
    """k_pixelinfo_tma (cp.async.bulk.tensor.2d + mbarrier, csrc/image_tma.cu) against k_pixelinfo -- the kernel the GPU
    suite pins against the reference's calculate_pixelinfo -- on a pseudo-random plane, including partial tiles."""
    import ctypes as C
    from dsopp_b200 import capi
    lib = capi.load_library()
    ms = (C.c_double * 2)()
    bad = C.c_int64(-1)
    rc = lib.dpba_debug_pixelinfo_ab(shape[0], shape[1], 3, ms, C.addressof(bad))
    assert rc == 0 and bad.value == 0, (rc, bad.value)
