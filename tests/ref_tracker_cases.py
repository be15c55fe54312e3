"""Windows for the tracker-side pins (tests/test_reference_tracker.py, tools/make_ref_tracker_golden.py): the coarse
tracker's reference depth maps (create_depth_maps.cpp) and the immature-landmark refine (landmarks_activator.cpp:122-316)."""
import numpy as np

from dsopp_b200 import synth
from oracle import activation_oracle as A
from oracle import features_oracle as FO
from oracle import pba_oracle as O

DEPTH_LEVELS = 4


def depth_case(seed=8):
    """-> (win, oracle frames, per-frame variances).  The window carries every skip rule of fillFineDepthMap: connection
    statuses other than kOk, outlier and marginalised landmarks, idepths below 1e-8 and negative ones."""
    win = synth.make_window(n_frames=5, points_per_frame=400, seed=seed, ab_scale=0.0)
    rng = np.random.default_rng(seed)
    n = len(win.frames)
    for r in range(n - 1):
        f = win.frames[r]
        M = len(f.idepth)
        win.statuses[(r, n - 1)][:] = rng.choice([0, 0, 0, 0, 0, 1, 3, 4], M)
        f.flags[rng.random(M) < 0.05] |= synth.FLAG_OUTLIER
        f.flags[rng.random(M) < 0.05] |= synth.FLAG_MARGINALIZED
        f.idepth[5], f.idepth[6] = 3e-9, -0.02
    frames = O.frames_from_window(win)
    variances = [rng.uniform(1e-7, 1e-3, len(f.idepth)) for f in win.frames[:-1]]
    return win, frames, variances


def track_landmarks(frames, variances):
    """What the TRACK holds when createReferenceDepthMaps reads it: the solver's landmarks after updateFrame's
    post-processing (photometric_bundle_adjustment.cpp:232-238: |idepth| < 1e-8 -> 0, other negative idepths -> outlier)."""
    tgt = frames[-1]
    out = []
    for k, f in enumerate(frames[:-1]):
        rho = np.array(f.idepth, dtype=np.float64)
        outlier = np.array(f.lm_outlier, dtype=bool)
        rho[np.abs(rho) < 1e-8] = 0.0
        outlier |= rho < 0
        var = np.full(len(rho), 1e-5) if variances is None else variances[k]
        out.append(dict(uv=f.uv, idepth=rho, idepth_variance=var, outlier=outlier, marginalized=f.lm_marginalized,
                        status=f.residuals[tgt.id].status))
    return out


def activation_case(seed=23):
    """-> (win, oracle ActFrames, raw images (n, H, W) float64, masks (n, H, W) uint8, candidates).  A candidate is
    (reference frame, landmark, starting idepth, minimum_inliers, sigma_huber); the list holds refinable, masked,
    out-of-view, negative and far-too-large starting depths."""
    win = synth.make_window(n_frames=5, points_per_frame=150, seed=seed, pose_noise=0.0, eps_scale=0.0, ab_scale=0.0)
    win.frames[2].mask[100:160, 200:330] = 0  # a masked region in one target
    images = np.stack([f.image[..., 0] for f in win.frames]).astype(np.float64)
    masks = np.stack([np.asarray(f.mask, dtype=np.uint8) for f in win.frames])
    # the reference builds {I, dx, dy} in double from the raw intensities: the oracle's frames get the same
    frames = [A.ActFrame(f.frame_id, f.T_w_lin, f.exposure, f.ab0, f.intr, FO.pixel_info(images[k]), f.mask)
              for k, f in enumerate(win.frames)]
    rng = np.random.default_rng(seed + 1)
    cands = []
    for r in (0, 3):
        f = win.frames[r]
        rho0 = (f.idepth_true * (1 + rng.uniform(-1, 1, len(f.idepth_true)) * 0.03)).astype(np.float32)
        rho0[:3] = [5.0, -0.5, 2000.0]  # hopeless / invalid candidates: must be deleted
        rho0[3:40] = f.idepth_true[3:40] * (1 + rng.uniform(-1, 1, 37) * 0.25)  # far from the optimum: rejected steps
        for l in range(len(rho0)):
            cands.append((r, l, float(rho0[l]), 3, 20.0))
        for l in range(0, 60, 2):
            cands.append((r, l, float(rho0[l]), 1, 2.0))
    return win, frames, images, masks, cands
