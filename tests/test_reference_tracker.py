"""The tracker-side oracles pinned against the REFERENCE'S OWN code.

oracle/build_ref_tracker.py compiles, from /root/reference and unchanged, the whole of create_depth_maps.cpp
(fillFineDepthMap, fillCoarseDepthMaps, dilateDepthMaps, createReferenceDepthMaps) and lines 122-316 of
landmarks_activator.cpp (class LandmarkActivationProblem, optimizeImmatureLandmark under the reference's LM driver).  The
track subsystem's containers they read are plain records (oracle/ref_stubs_track); all arithmetic is the reference's.
tests/golden/ref_tracker.npz holds their outputs on the windows of tests/ref_tracker_cases.py.

  * oracle/depth_map_oracle.py: the same pixels filled on every level (exact), weights to 1e-12, idepth sums to 1e-12
    (the depth scale is one dot product summed in a different order);
  * oracle/activation_oracle.py: every activate / delete decision (exact), refined inverse depths to 1e-9;
  * the CUDA path against the same golden vectors directly, at its fp32 bars (tests/test_gpu_activation.py,
    tests/test_gpu_device_paths.py state them).
"""
import os

import numpy as np
import pytest

import ref_tracker_cases as TC
from oracle import activation_oracle as A
from oracle import depth_map_oracle as D
from oracle import ref_tracker

GOLDEN_PATH = os.path.join(os.path.dirname(__file__), "golden", "ref_tracker.npz")
needs_ref = pytest.mark.skipif(not ref_tracker.available(), reason="neither /root/reference nor oracle/_ref is present")


@pytest.fixture(scope="module")
def golden():
    g = np.load(GOLDEN_PATH)
    return {k: g[k] for k in g.files}


def _maps_equal(got, ref_i, ref_w, level):
    gi, gw = got
    assert gi.shape == ref_i.shape
    assert np.array_equal(gw > 0, ref_w > 0), level                      # the same pixels, after the dilation too
    assert np.abs(gw - ref_w).max() <= 1e-12 * ref_w.max(), level
    assert np.abs(gi - ref_i).max() <= 1e-12 * np.abs(ref_i).max(), level


@pytest.mark.parametrize("tag", ["const", "var"])
def test_depth_map_oracle_reproduces_the_reference_golden(golden, tag):
    _, frames, variances = TC.depth_case()
    maps = D.create_reference_depth_maps(frames, TC.DEPTH_LEVELS, None if tag == "const" else variances)
    for l, m in enumerate(maps):
        _maps_equal(m, golden[f"depth/{tag}/idepth{l}"], golden[f"depth/{tag}/weight{l}"], l)
    # the case reaches the skip rules and the dilation
    w0, w3 = golden[f"depth/{tag}/weight0"], golden[f"depth/{tag}/weight3"]
    n_candidates = sum(len(f.idepth) for f in frames[:-1])
    assert 0 < (w0 > 0).sum() and (w3 > 0).sum() > 1000
    _, fine_w = D.fill_fine_depth_map(frames, None if tag == "const" else variances)
    assert (fine_w > 0).sum() < 0.8 * n_candidates                      # statuses / outliers / marginalised skipped
    assert (w0 > 0).sum() > 2 * (fine_w > 0).sum()                      # empty neighbours filled by dilateDepthMaps


def test_activation_oracle_reproduces_the_reference_golden(golden):
    win, frames, _, _, cands = TC.activation_case()
    status, idepth = golden["activation/status"], golden["activation/idepth"]
    assert len(cands) == len(status)
    worst = 0.0
    for (r, l, rho0, min_inl, sigma), s_ref, rho_ref in zip(cands, status, idepth):
        f = win.frames[r]
        act, rho, _ = A.optimize_immature_landmark(frames[r], frames, f.uv[l], f.patch[l], rho0, min_inl, sigma)
        assert act == (s_ref == 0), (r, l, rho0)                         # kActivate = 0, kDelete = 2
        if act:
            worst = max(worst, abs(rho - rho_ref) / abs(rho_ref))
        else:
            assert rho_ref == np.float64(rho0)                           # a deleted landmark keeps its interval
    assert worst <= 1e-9, worst
    assert (status == 2).sum() >= 10 and (status == 0).sum() >= 300     # both outcomes


@needs_ref
def test_golden_is_what_the_reference_computes(golden):
    _, frames, variances = TC.depth_case()
    tgt = frames[-1]
    maps = ref_tracker.create_reference_depth_maps([f.t_world_agent() for f in frames], tgt.intr, tgt.W, tgt.H,
                                                   TC.DEPTH_LEVELS, TC.track_landmarks(frames, variances))
    for l, (idw, wgt) in enumerate(maps):
        assert np.array_equal(idw, golden[f"depth/var/idepth{l}"]) and np.array_equal(wgt, golden[f"depth/var/weight{l}"])
    win, _, images, masks, cands = TC.activation_case()
    T, e, ab = [f.T_w_lin for f in win.frames], [f.exposure for f in win.frames], [f.ab0 for f in win.frames]
    for k in range(0, len(cands), 9):
        r, l, rho0, min_inl, sigma = cands[k]
        f = win.frames[r]
        s, rho = ref_tracker.optimize_immature_landmark(T, e, ab, images, masks, f.intr, r, f.uv[l], f.patch[l], rho0, rho0,
                                                        min_inl, sigma)
        assert s == golden["activation/status"][k] and rho == golden["activation/idepth"][k]


@needs_ref
def test_other_windows_live():
    """The pin does not hang on the seed of the golden file."""
    for seed in (31, 32):
        _, frames, variances = TC.depth_case(seed)
        tgt = frames[-1]
        ref = ref_tracker.create_reference_depth_maps([f.t_world_agent() for f in frames], tgt.intr, tgt.W, tgt.H,
                                                      TC.DEPTH_LEVELS, TC.track_landmarks(frames, variances))
        for l, (m, (ri, rw)) in enumerate(zip(D.create_reference_depth_maps(frames, TC.DEPTH_LEVELS, variances), ref)):
            _maps_equal(m, ri, rw, l)
        win, aframes, images, masks, cands = TC.activation_case(seed)
        T, e, ab = [f.T_w_lin for f in win.frames], [f.exposure for f in win.frames], [f.ab0 for f in win.frames]
        for r, l, rho0, min_inl, sigma in cands[::5]:
            f = win.frames[r]
            s, rho_ref = ref_tracker.optimize_immature_landmark(T, e, ab, images, masks, f.intr, r, f.uv[l], f.patch[l], rho0,
                                                                rho0, min_inl, sigma)
            act, rho, _ = A.optimize_immature_landmark(aframes[r], aframes, f.uv[l], f.patch[l], rho0, min_inl, sigma)
            assert act == (s == 0) and (not act or abs(rho - rho_ref) <= 1e-9 * abs(rho_ref)), (seed, r, l)


@pytest.mark.gpu
def test_device_depth_maps_against_the_reference_golden(golden):
    """dpba_create_reference_depth_maps against the reference's maps: fp32 reprojection may move a landmark that sits
    within ~1e-4 px of a rounding boundary to the neighbouring pixel (bounded count); values 1e-4."""
    from dsopp_b200 import capi
    win, _, _ = TC.depth_case()
    h = capi.upload_window(win)
    h.first_estimate()
    got = h.create_reference_depth_maps(TC.DEPTH_LEVELS, 1e-5)
    for lvl, (gi, gw) in enumerate(got):
        ri, rw = golden[f"depth/const/idepth{lvl}"], golden[f"depth/const/weight{lvl}"]
        g_on, r_on = gw > 0, rw > 0
        assert (g_on != r_on).sum() <= 12 * (1 + (lvl == 0)), (lvl, (g_on != r_on).sum())
        both = g_on & r_on
        assert (np.abs(gw[both] - rw[both]) > 1e-4 * np.abs(rw[both])).sum() <= 12, lvl
        assert (np.abs(gi[both] - ri[both]) > 1e-4 * np.abs(ri[both])).sum() <= 12, lvl
    h.close()


@pytest.mark.gpu
def test_device_activation_refine_against_the_reference_golden(golden):
    """dpba_refine_immature_landmarks against optimizeImmatureLandmark's decisions and inverse depths (fp32: 95 % within 2e-4,
    all within 1e-3; decisions exact bar <= 1 % on fp32 noise of a boundary)."""
    from dsopp_b200 import capi
    win, _, _, _, cands = TC.activation_case()
    h = capi.upload_window(win)
    status, idepth = golden["activation/status"], golden["activation/idepth"]
    total = flips = 0
    rel = []
    for key in sorted({(r, m, s) for r, _, _, m, s in cands}):
        idx = [k for k, c in enumerate(cands) if (c[0], c[3], c[4]) == key]
        r, min_inl, sigma = key
        f = win.frames[r]
        ls = np.array([cands[k][1] for k in idx])
        rho0 = np.array([cands[k][2] for k in idx], dtype=np.float32)
        got_rho, got_act, _ = h.refine_immature_landmarks(r, f.uv[ls], rho0, f.patch[ls], min_inl, sigma)
        for j, k in enumerate(idx):
            total += 1
            if bool(got_act[j]) != (status[k] == 0):
                flips += 1
            elif got_act[j]:
                rel.append(abs(got_rho[j] - idepth[k]) / abs(idepth[k]))
    rel = np.array(rel)
    print(f"[activation vs reference] {total} candidates, {flips} flips, rel err 95% {np.quantile(rel, 0.95):.1e} max {rel.max():.1e}")
    assert flips <= max(1, total // 100)
    assert np.quantile(rel, 0.95) <= 2e-4 and rel.max() <= 1e-3
    h.close()
