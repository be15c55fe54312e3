"""Properties of oracle/depth_map_oracle.py (createReferenceDepthMaps, src/tracker/tracker/src/create_depth_maps.cpp:19-146).
The reference holds no test or fixture for this function (test/test/tracker/reference_frame/
test_reference_frame_depth_map.cpp only checks the consumer side, LocalFrame's depth-map constructor); the value pin against
the reference's own compiled file is tests/test_reference_tracker.py."""
import numpy as np
import pytest

from dsopp_b200 import synth
from oracle import depth_map_oracle as D
from oracle import pba_oracle as O


def exact_window(n_frames=4, pts=300, seed=2):
    """Estimated state == rendering state, so that a splatted inverse depth can be compared with the scene's."""
    win = synth.make_window(n_frames=n_frames, points_per_frame=pts, seed=seed, pose_noise=0.0, idepth_noise=0.0,
                            eps_scale=0.0, ab_scale=0.0)
    return win, O.frames_from_window(win)


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_vectorised_dilation_equals_the_reference_loops(level):
    rng = np.random.default_rng(level)
    H, W = 23, 31
    wgt = np.where(rng.random((H, W)) < 0.25, rng.uniform(0.5, 20.0, (H, W)), 0.0)
    idw = wgt * rng.uniform(0.05, 0.5, (H, W))
    a = D.dilate(idw, wgt, level)
    b = D.dilate_loops(idw, wgt, level)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # only empty interior pixels change; levels 0-1 use the diagonal neighbours, levels >= 2 the axis neighbours
    changed = a[1] != wgt
    assert not changed[wgt > 0].any()
    assert not changed[0].any() and not changed[-1].any() and not changed[:, 0].any() and not changed[:, -1].any()
    y, x = np.argwhere(changed)[0]
    nb = [(1, 0), (-1, 0), (0, 1), (0, -1)] if level > 1 else [(1, 1), (-1, -1), (1, -1), (-1, 1)]
    assert any(wgt[y + dy, x + dx] > 0 for dx, dy in nb)


def test_coarse_levels_are_2x2_sums():
    rng = np.random.default_rng(1)
    idw, wgt = rng.random((48, 64)), rng.random((48, 64))
    i1, w1 = D.fill_coarse(idw, wgt)
    assert i1.shape == (24, 32)
    assert np.isclose(w1.sum(), wgt.sum()) and np.isclose(i1.sum(), idw.sum())
    assert np.isclose(w1[3, 5], wgt[6:8, 10:12].sum())
    # odd sizes: the last row / column is dropped (integer halving of the pyramid)
    i2, w2 = D.fill_coarse(idw[:47, :63], wgt[:47, :63])
    assert w2.shape == (23, 31)


def test_constant_variance_weights_and_exclusions():
    win, frames = exact_window()
    tgt = frames[-1]
    idw, wgt = D.fill_fine_depth_map(frames)
    w_each = np.sqrt(1e-3 / (1e-5 + 1e-12))   # create_depth_maps.cpp:52 with the constant of photometric_bundle_adjustment.cpp:254
    n_splat = int(round(wgt.sum() / w_each))
    assert np.isclose(wgt.sum(), n_splat * w_each)
    assert 0 < n_splat <= sum(len(f.idepth) for f in frames[:-1])
    # the newest keyframe's own landmarks are not splatted (loop to frames.size() - 1, :27)
    assert n_splat <= sum(len(f.idepth) for f in frames[:-1])
    # statuses other than kOk, outliers and marginalised landmarks are skipped (:36-40)
    frames[0].residuals[tgt.id].status[:] = O.K_OUTLIER
    frames[1].lm_outlier[:] = True
    frames[2].lm_marginalized[:] = True
    _, w0 = D.fill_fine_depth_map(frames)
    assert w0.sum() == 0.0


def test_splatted_inverse_depth_is_the_scene_seen_from_the_newest_keyframe():
    win, frames = exact_window(n_frames=5, pts=400, seed=3)
    sf = win.frames[-1]
    depth = synth.plane_depth(sf.T_w_true, sf.intr, win.width, win.height)
    idw, wgt = D.fill_fine_depth_map(frames)
    filled = wgt > 0
    assert filled.sum() > 500
    rho = idw[filled] / wgt[filled]
    # the splat lands on the ROUNDED pixel: half a pixel of a 5 m plane changes 1/z by < 1e-3 relative
    assert np.abs(rho * depth[filled] - 1.0).max() < 2e-3
    # inside the ROI of the target only (scalar reproject success, camera_reproject.hpp:286-290)
    ys, xs = np.nonzero(filled)
    assert xs.min() >= 4 and ys.min() >= 4 and xs.max() <= win.width - 5 and ys.max() <= win.height - 5


def test_track_post_processing_of_inverse_depths():
    """updateFrame (photometric_bundle_adjustment.cpp:232-238) stands between the solver and the track the depth maps read:
    negative inverse depths are outliers there, tiny ones are zero."""
    win, frames = exact_window(n_frames=3, pts=60, seed=7)
    base = D.fill_fine_depth_map(frames)[1].sum()
    frames[0].lm_outlier[:10] = True
    without = D.fill_fine_depth_map(frames)[1].sum()       # the map without the first ten landmarks of frame 0
    frames[0].lm_outlier[:10] = False
    assert without < base
    frames[0].idepth[:10] = -0.02
    assert np.isclose(D.fill_fine_depth_map(frames)[1].sum(), without)   # negative: outliers in the track, not splatted
    frames[0].idepth[:10] = 5e-9
    idw, wgt = D.fill_fine_depth_map(frames)
    assert wgt.sum() > without                  # tiny: they count again (where the reprojection stays in the image) ...
    assert np.all(idw >= 0)                     # ... with inverse depth exactly 0


def test_uncertainty_weights():
    win, frames = exact_window(n_frames=3, pts=100, seed=4)
    var = [np.full(len(f.idepth), 1e-5) for f in frames]
    var[0][:] = 4e-5
    a = D.fill_fine_depth_map(frames)
    b = D.fill_fine_depth_map(frames, var)
    # quartering the precision of frame 0's landmarks halves their weights; the weighted MEAN is what consumers read
    assert b[1].sum() < a[1].sum()
    both = (a[1] > 0)
    assert np.array_equal(both, b[1] > 0)
    single = both & (np.isclose(a[1], np.sqrt(1e-3 / (1e-5 + 1e-12))))
    assert np.allclose(a[0][single] / a[1][single], b[0][single] / b[1][single], rtol=1e-12)


def test_pyramid_of_maps():
    win, frames = exact_window(n_frames=4, pts=300, seed=5)
    maps = D.create_reference_depth_maps(frames, 4)
    assert [m[1].shape for m in maps] == [(480, 640), (240, 320), (120, 160), (60, 80)]
    for idw, wgt in maps:
        ok = wgt > 0
        assert ok.any() and np.all(idw[ok] > 0)
    # dilation only adds pixels: every level has more non-empty pixels than the sums alone
    fine = D.fill_fine_depth_map(frames)
    assert (maps[0][1] > 0).sum() > (fine[1] > 0).sum()
    # mean inverse depth is preserved by the 2x2 sums (weighted mean of weighted means)
    l1 = D.fill_coarse(*fine)
    ok = l1[1] > 0
    blocks = fine[0].reshape(240, 2, 320, 2).sum(axis=(1, 3))
    assert np.allclose(l1[0][ok], blocks[ok])
