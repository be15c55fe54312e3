"""CPU oracle (test infrastructure only) of the coarse tracker's reference depth maps -- the producer of the input of
the pose aligner ("depth-map build create_depth_maps.cpp:19-147").

Restates, in NumPy float64 (the reference's default Precision):
  createReferenceDepthMaps   src/tracker/tracker/src/create_depth_maps.cpp:122-146
    fillFineDepthMap         :19-58   splat every kOk, non-outlier, non-marginalised landmark of the older keyframes into
                                       the NEWEST keyframe, weighted by sqrt(1e-3 / (idepth variance + 1e-12))
    fillCoarseDepthMaps      :70-88   level l = 2x2 SUM (idepth*weight and weight alike) of level l-1
    dilateDepthMaps          :90-120  empty pixels take the mean of their non-empty neighbours (diagonal neighbours on
                                       levels 0-1, axis neighbours on levels >= 2), borders excluded
  DepthMap::WeightedIdepth   src/energy/problems/include/energy/problems/depth_map.hpp:17-31
  getDepthScale              src/energy/camera_model/include/energy/camera_model/camera_model_base.hpp:103-107
  scalar pinhole reproject   src/energy/projector/include/energy/projector/camera_reproject.hpp:270-293
  call site                  src/tracker/tracker/src/monocular_tracker.cpp:465,509 (right after the BA solve)

PARITY PINNED: create_depth_maps.cpp compiles here WHOLE and unchanged (oracle/build_ref_tracker.py; only the track
containers it reads are stand-in records) and this restatement fills the same pixels on every level, dilation included, with
weights and inverse-depth sums at 1e-12 (tests/test_reference_tracker.py, tests/golden/ref_tracker.npz); properties in
tests/test_depth_map_oracle.py.

Layout: a map is a pair of (H, W) arrays (idepth_w, weight); the reference's `map(x, y)` is element [y, x].
`idepth_w` is the WEIGHTED SUM of inverse depths, as in the reference: consumers divide by `weight`
(local_frame.hpp:379).
"""
import numpy as np

from . import pba_oracle as O

K_EPS = 1e-12              # create_depth_maps.cpp:23
K_VARIATION_SCALE = 1e-3   # :25


def _round_half_away(x):
    """Eigen's array round() == std::round: halves away from zero (:46)."""
    return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5)).astype(np.int64)


def fill_fine_depth_map(frames, idepth_variance=None):
    """:19-58.  frames: oracle Frames in window order, the last one is the target.  idepth_variance: list of (M_f,) arrays
    (ActiveTrackingLandmark::idepthVariance, = inv_hessian_idepth_idepth after a solve with uncertainty,
    photometric_bundle_adjustment.cpp:252) or None for the constant 1e-5 of :254."""
    tgt = frames[-1]
    idw = np.zeros((tgt.H, tgt.W))
    wgt = np.zeros((tgt.H, tgt.W))
    for fi, ref in enumerate(frames[:-1]):
        T_t_r = O.se3_inv(tgt.t_world_agent()) @ ref.t_world_agent()
        rp = O.Reprojector(ref, tgt, T_t_r)
        status = ref.residuals[tgt.id].status
        A = rp.reproject_
        for l in range(len(ref.idepth)):
            if status[l] != O.K_OK:
                continue
            if ref.lm_outlier[l] or ref.lm_marginalized[l]:
                continue
            rho = ref.idepth[l]
            # the reference reads the TRACK's landmarks: the solver's after updateFrame's post-processing
            # (photometric_bundle_adjustment.cpp:232-238) -- |idepth| < 1e-8 -> 0, other negative idepths -> outlier
            if abs(rho) < 1e-8:
                rho = 0.0
            elif rho < 0:
                continue
            uv = ref.uv[l]
            # scalar reproject, camera_reproject.hpp:270-293 (one point: ROI of that point only)
            ok = bool(O.valid_idepth(rho)) and _in_roi(uv, ref.W, ref.H)
            p = A[:, :2] @ uv + A[:, 2] + A[:, 3] * rho
            if not (p[2] > 0):
                continue
            t2 = p[:2] / p[2]
            ok = ok and _in_roi(t2, tgt.W, tgt.H)
            if not ok:
                continue
            ix, iy = _round_half_away(t2)
            fx, fy, cx, cy = ref.intr
            direction = np.array([(uv[0] - cx) / fx, (uv[1] - cy) / fy, 1.0])  # pinhole unproject, pinhole_camera.hpp:137-139
            depth_scale = (T_t_r[:3, :3] @ direction + T_t_r[:3, 3] * rho)[2]      # getDepthScale
            var = 1e-5 if idepth_variance is None else float(idepth_variance[fi][l])
            w = np.sqrt(K_VARIATION_SCALE / (var + K_EPS))
            idw[iy, ix] += rho / depth_scale * w
            wgt[iy, ix] += w
    return idw, wgt


def _in_roi(p, W, H):
    return bool(p[0] >= O.BORDER and p[1] >= O.BORDER and p[0] <= W - O.BORDER - 1 and p[1] <= H - O.BORDER - 1)


def fill_coarse(idw, wgt):
    """One level of fillCoarseDepthMaps (:70-88): width/height halve with integer division, 2x2 sums."""
    H2, W2 = idw.shape[0] // 2, idw.shape[1] // 2

    def s(a):
        return a[0:2 * H2:2, 0:2 * W2:2] + a[0:2 * H2:2, 1:2 * W2:2] + a[1:2 * H2:2, 0:2 * W2:2] + a[1:2 * H2:2, 1:2 * W2:2]

    return s(idw), s(wgt)


def dilate(idw, wgt, level):
    """One level of dilateDepthMaps (:90-120).  Only pixels that were empty (weight <= 0) change and only non-empty
    neighbours are read, so the in-place update of the reference is a pure function of the input."""
    H, W = wgt.shape
    off = [(1, 0), (-1, 0), (0, 1), (0, -1)] if level > 1 else [(1, 1), (-1, -1), (1, -1), (-1, 1)]  # (dx, dy), :103-107
    out_i, out_w = idw.copy(), wgt.copy()
    if H < 3 or W < 3:
        return out_i, out_w
    s = np.zeros((H - 2, W - 2))
    num = np.zeros((H - 2, W - 2))
    numn = np.zeros((H - 2, W - 2))
    for dx, dy in off:
        nb_w = wgt[1 + dy:H - 1 + dy, 1 + dx:W - 1 + dx]
        nb_i = idw[1 + dy:H - 1 + dy, 1 + dx:W - 1 + dx]
        has = nb_w > 0
        s += np.where(has, nb_i, 0.0)
        num += np.where(has, nb_w, 0.0)
        numn += has
    fill = (wgt[1:H - 1, 1:W - 1] <= 0) & (numn > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        out_i[1:H - 1, 1:W - 1] = np.where(fill, s / numn, idw[1:H - 1, 1:W - 1])
        out_w[1:H - 1, 1:W - 1] = np.where(fill, num / numn, wgt[1:H - 1, 1:W - 1])
    return out_i, out_w


def create_reference_depth_maps(frames, n_levels, idepth_variance=None):
    """createReferenceDepthMaps (:122-146) -> list over levels of (idepth_w, weight), each (H_l, W_l)."""
    maps = [fill_fine_depth_map(frames, idepth_variance)]
    for _ in range(1, n_levels):
        maps.append(fill_coarse(*maps[-1]))
    return [dilate(i, w, lvl) for lvl, (i, w) in enumerate(maps)]


def dilate_loops(idw, wgt, level):
    """The same function written as the reference's loops (small maps only): checks the vectorised form above."""
    H, W = wgt.shape
    backup = wgt.copy()
    oi, ow = idw.copy(), wgt.copy()
    off = [(1, 0), (-1, 0), (0, 1), (0, -1)] if level > 1 else [(1, 1), (-1, -1), (1, -1), (-1, 1)]
    for y in range(1, H - 1):
        for x in range(1, W - 1):
            if backup[y, x] <= 0:
                s = n = nn = 0.0
                for dx, dy in off:
                    if backup[y + dy, x + dx] > 0:
                        s += oi[y + dy, x + dx]
                        n += backup[y + dy, x + dx]
                        nn += 1
                if nn > 0:
                    oi[y, x] = s / nn
                    ow[y, x] = n / nn
    return oi, ow
