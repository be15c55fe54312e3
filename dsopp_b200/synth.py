"""Seeded synthetic sliding windows for the photometric bundle-adjustment hot path.

The reference's own fixture (test/test_data/track30seconds, built by
test/tools/src/solver_test_data.cpp:31-143) is not in the checkout, so every
parity / bench input is generated here: a textured,
tilted plane seen by N pinhole keyframes, rendered analytically by ray/plane
intersection, with per-frame exposure and affine brightness applied
consistently so the true minimum of the photometric energy exists.

This module is product-side input plumbing (bench.py, tests, smoke); it does
not import anything from oracle/.

Conventions follow the reference:
  * image grid = interleaved {I, dx, dy} per pixel (features/camera/pixel_map.hpp:126-131),
    gradients by central differences, one-sided (x1.0) on the border
    (features/src/calculate_pixelinfo.cpp:340-374);
  * 8-point DSO pattern (common/pattern/pattern.hpp:15-35);
  * pose state = T_w_lin * exp(eps), tangent order [translation(3), rotation(3)]
    (Sophus convention, energy/motion/se3_motion.hpp:231-239);
  * statuses: PointConnectionStatus{kOk=0,kOutlier,kOccluded,kOOB,kUnknown}
    (track/connections/frame_connection.hpp:19-25).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

# common/pattern/pattern.hpp:22-33 -- (x_i, y_i) pairs, centre index 4
PATTERN = np.array(
    [[0, 2], [-1, 1], [1, 1], [-2, 0], [0, 0], [2, 0], [-1, -1], [0, -2]], dtype=np.float64
)
PATTERN_SIZE = 8
BLOCK = 8  # Motion::DoF + 2

K_OK, K_OUTLIER, K_OCCLUDED, K_OOB, K_UNKNOWN = 0, 1, 2, 3, 4

FLAG_MARGINALIZED = 1
FLAG_TO_MARGINALIZE = 2
FLAG_OUTLIER = 4
FLAG_ILL_CONDITIONED = 8


def hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def se3_exp(xi):
    """Sophus::SE3::exp restated (tangent = [upsilon; omega]); returns 4x4.

    R = Exp(omega), t = V(omega) upsilon,
    V = I + (1-cos th)/th^2 W + (th - sin th)/th^3 W^2   (SURVEY.md section 8c).
    """
    xi = np.asarray(xi, dtype=np.float64)
    v, w = xi[:3], xi[3:]
    th2 = float(w @ w)
    th = np.sqrt(th2)
    W = hat(w)
    if th < 1e-10:
        R = np.eye(3) + W + 0.5 * W @ W
        V = np.eye(3) + 0.5 * W + (1.0 / 6.0) * W @ W
    else:
        a = np.sin(th) / th
        b = (1.0 - np.cos(th)) / th2
        c = (th - np.sin(th)) / (th2 * th)
        R = np.eye(3) + a * W + b * W @ W
        V = np.eye(3) + b * W + c * W @ W
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ v
    return T


def se3_inv(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def pixelinfo(I: np.ndarray) -> np.ndarray:
    """{I,dx,dy} packing, calculate_pixelinfo.cpp:340-374 (scalar definition), same dtype as I."""
    I = np.asarray(I)
    H, W = I.shape
    out = np.empty((H, W, 3), dtype=I.dtype)
    half = I.dtype.type(0.5)
    out[..., 0] = I
    out[:, 1:-1, 1] = half * (I[:, 2:] - I[:, :-2])
    out[:, 0, 1] = I[:, 1] - I[:, 0]
    out[:, -1, 1] = I[:, -1] - I[:, -2]
    out[1:-1, :, 2] = half * (I[2:, :] - I[:-2, :])
    out[0, :, 2] = I[1, :] - I[0, :]
    out[-1, :, 2] = I[-1, :] - I[-2, :]
    return out


def downscale(I: np.ndarray) -> np.ndarray:
    """2x2 box filter, features/internal/features/camera/downscale_image.hpp:16-33."""
    q = I.dtype.type(0.25)
    H, W = I.shape
    H2, W2 = H // 2, W // 2
    a = I[0 : 2 * H2 : 2, 0 : 2 * W2 : 2]
    b = I[1 : 2 * H2 : 2, 1 : 2 * W2 : 2]
    c = I[0 : 2 * H2 : 2, 1 : 2 * W2 : 2]
    d = I[1 : 2 * H2 : 2, 0 : 2 * W2 : 2]
    return q * (a + b + c + d)


@dataclass
class SynthFrame:
    frame_id: int
    timestamp: int
    T_w_lin: np.ndarray  # 4x4 float64 linearisation point
    T_w_true: np.ndarray  # 4x4 float64, pose used for rendering
    exposure: float
    ab0: np.ndarray  # (2,) affine brightness (a, b)
    intr: np.ndarray  # (4,) fx, fy, cx, cy
    image: np.ndarray  # (H, W, 3) float32 {I,dx,dy}
    mask: np.ndarray  # (H, W) uint8
    fixed: bool
    state_eps: np.ndarray  # (8,)
    # landmarks hosted by this frame
    uv: np.ndarray  # (M,2) float64 (integer valued)
    idepth: np.ndarray  # (M,)
    idepth_true: np.ndarray  # (M,)
    patch: np.ndarray  # (M,8)
    flags: np.ndarray  # (M,) uint8
    pyramid: List[np.ndarray] = field(default_factory=list)  # levels 1.. of {I,dx,dy}


@dataclass
class SynthWindow:
    frames: List[SynthFrame]
    width: int
    height: int
    statuses: dict  # (ref_idx, tgt_idx) -> (M_ref,) uint8
    seed: int

    @property
    def n_frames(self):
        return len(self.frames)

    @property
    def units(self):
        n = self.n_frames
        return sum(len(f.idepth) for f in self.frames) * (n - 1)


def _render(T_w_c, intr, W, H, n, d, e1, e2, tex):
    fx, fy, cx, cy = intr
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dirs_c = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)
    R, t = T_w_c[:3, :3], T_w_c[:3, 3]
    dirs_w = dirs_c @ R.T
    lam = (d - n @ t) / (dirs_w @ n)  # depth along camera z (dirs_c.z == 1)
    X = t + lam[..., None] * dirs_w
    return tex(X @ e1, X @ e2), lam


def make_window(
    n_frames: int = 8,
    points_per_frame: int = 2000,
    width: int = 640,
    height: int = 480,
    seed: int = 0,
    pose_noise: float = 1e-3,
    idepth_noise: float = 2e-3,
    eps_scale: float = 1e-3,
    levels: int = 1,
    marginalize_first: bool = False,
    ab_scale: float = 1.0,
) -> SynthWindow:
    """Build the SURVEY.md section 8(d) window.

    pose_noise perturbs the *estimated* pose away from the rendering pose (so the GN step is
    non-trivial); eps_scale is the test_linear_system.cpp:101-103 style state_eps with the
    linearisation point pre-compensated so that T_lin * exp(eps) == estimated pose.
    """
    rng = np.random.default_rng(seed)
    W, H = width, height
    f0 = 400.0 * W / 640.0
    intr = np.array([f0, f0, W / 2.0, H / 2.0])

    n = np.array([0.1, -0.05, 1.0])
    n /= np.linalg.norm(n)
    Z0 = 5.0
    d = n @ np.array([0.0, 0.0, Z0])
    e1 = np.cross(n, [0.0, 1.0, 0.0])
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(n, e1)

    nk = 12
    amp = rng.uniform(4.0, 20.0, nk)
    kmag = rng.uniform(2.0, 40.0, nk)
    kang = rng.uniform(0.0, 2 * np.pi, nk)
    kvec = np.stack([kmag * np.cos(kang), kmag * np.sin(kang)], axis=1)
    phase = rng.uniform(0.0, 2 * np.pi, nk)

    def tex(x, y):
        val = np.full_like(x, 128.0)
        for a, kv, ph in zip(amp, kvec, phase):
            val += a * np.sin(kv[0] * x + kv[1] * y + ph)
        return np.clip(val, 0.0, 255.0)

    frames: List[SynthFrame] = []
    for k in range(n_frames):
        trans = k * np.array([0.06, 0.01, 0.03]) + rng.uniform(-1, 1, 3) * 0.01
        rot = rng.uniform(-1, 1, 3) * 0.02
        T_true = se3_exp(np.concatenate([np.zeros(3), rot]))
        T_true[:3, 3] = trans
        tau = rng.uniform(0.8, 1.2)
        # ab_scale = 0 is the production-like case: fabric.cpp:68-69 pins (a, b) with regularisers
        # (1e12, 1e8), i.e. photometrically calibrated input whose true affine brightness is zero.
        a = rng.uniform(-0.05, 0.05) * ab_scale
        b = rng.uniform(-3.0, 3.0) * ab_scale
        rad, depth = _render(T_true, intr, W, H, n, d, e1, e2, tex)
        I = (tau * np.exp(a) * rad + b).astype(np.float32)
        img = pixelinfo(I)
        pyr = []
        Il = I
        for _ in range(1, levels):
            Il = downscale(Il)
            pyr.append(pixelinfo(Il))

        # integer pixel coordinates without replacement in [12, W-13] x [12, H-13]
        nx, ny = W - 24, H - 24
        M = min(points_per_frame, nx * ny)
        flat = rng.choice(nx * ny, size=M, replace=False)
        px = (flat % nx + 12).astype(np.float64)
        py = (flat // nx + 12).astype(np.float64)
        uv = np.stack([px, py], axis=1)
        idepth_true = 1.0 / depth[py.astype(int), px.astype(int)]
        idepth = idepth_true + rng.uniform(-1, 1, M) * idepth_noise
        pi = (px[:, None] + PATTERN[None, :, 0]).astype(int)
        pj = (py[:, None] + PATTERN[None, :, 1]).astype(int)
        patch = I[pj, pi].astype(np.float64)

        # estimated pose = true pose * exp(noise); frame 0 stays exact (it is the fixed gauge)
        noise = rng.uniform(-1, 1, 6) * (pose_noise if k > 0 else 0.0)
        T_est = T_true @ se3_exp(noise)
        eps = np.zeros(8)
        # the fixed frame's prior pulls its eps to zero (problem.hpp:51-55), so it starts at zero
        eps[:6] = rng.uniform(-1, 1, 6) * (eps_scale if k > 0 else 0.0)
        T_lin = T_est @ se3_exp(-eps[:6])

        frames.append(
            SynthFrame(
                frame_id=k,
                timestamp=1000 * (k + 1),
                T_w_lin=T_lin,
                T_w_true=T_true,
                exposure=float(tau),
                ab0=np.array([a, b]),
                intr=intr.copy(),
                image=img,
                mask=np.full((H, W), 255, dtype=np.uint8),
                fixed=(k == 0),
                state_eps=eps,
                uv=uv,
                idepth=idepth,
                idepth_true=idepth_true,
                patch=patch,
                flags=np.zeros(M, dtype=np.uint8),
                pyramid=pyr,
            )
        )

    statuses = {}
    for r in range(n_frames):
        for t in range(n_frames):
            if r != t:
                statuses[(r, t)] = np.zeros(len(frames[r].idepth), dtype=np.uint8)

    if marginalize_first:
        # config 5: every landmark of KF 0 flagged to_marginalize (local_frame.hpp:493-496 sets
        # to_marginalize together with is_marginalized)
        frames[0].flags[:] = FLAG_MARGINALIZED | FLAG_TO_MARGINALIZE

    return SynthWindow(frames=frames, width=W, height=H, statuses=statuses, seed=seed)


def scene_plane():
    """The textured plane every synthetic frame looks at: unit normal n and offset d (n . X = d)."""
    n = np.array([0.1, -0.05, 1.0])
    n /= np.linalg.norm(n)
    return n, float(n @ np.array([0.0, 0.0, 5.0]))


def plane_depth(T_w_c, intr, W, H):
    """Per-pixel depth (along the camera z axis) of the scene plane seen from pose T_w_c, (H, W) float64."""
    n, d = scene_plane()
    fx, fy, cx, cy = intr
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dirs_c = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)
    R, t = T_w_c[:3, :3], T_w_c[:3, 3]
    return (d - n @ t) / ((dirs_c @ R.T) @ n)


@dataclass
class AlignmentCase:
    """Input of the coarse tracker's direct image alignment: the last keyframe with a
    reference depth map (reference frame, fixed) and a new frame with a perturbed pose guess (target frame)."""
    reference: SynthFrame
    target: SynthFrame
    T_w_target_guess: np.ndarray   # 4x4, what the tracker's motion model proposes
    T_t_r_true: np.ndarray         # 4x4 ground truth target <- reference
    idepth_sum: np.ndarray         # (H, W) depth-map accumulators as create_depth_maps.cpp leaves them
    weight: np.ndarray             # (H, W)
    width: int
    height: int


def make_alignment_case(seed=0, width=640, height=480, density=0.05, pose_noise=5e-3, idepth_noise=2e-3,
                        ab_scale=0.0) -> AlignmentCase:
    """density = fraction of pixels that carry depth (1.0 = BASELINE.json configs[2]'s synthetic full-frame bound; the
    reference's own depth map is sparse: splatted landmarks + one dilation, tracker/src/create_depth_maps.cpp:19-122)."""
    win = make_window(n_frames=2, points_per_frame=1, width=width, height=height, seed=seed, pose_noise=0.0, eps_scale=0.0,
                      ab_scale=ab_scale)
    ref, tgt = win.frames
    rng = np.random.default_rng(seed + 77)
    depth = plane_depth(ref.T_w_true, ref.intr, width, height)
    weight = (rng.random((height, width)) < density).astype(np.float64) * rng.integers(1, 4, (height, width))
    idepth = 1.0 / depth + rng.uniform(-1, 1, (height, width)) * idepth_noise
    guess = tgt.T_w_true @ se3_exp(rng.uniform(-1, 1, 6) * pose_noise)
    return AlignmentCase(reference=ref, target=tgt, T_w_target_guess=guess,
                         T_t_r_true=np.linalg.inv(tgt.T_w_true) @ ref.T_w_true, idepth_sum=idepth * weight, weight=weight,
                         width=width, height=height)


CONFIGS = {
    # BASELINE.json configs[0]: correctness anchor (replaces the absent track30seconds)
    "anchor": dict(n_frames=3, points_per_frame=200),
    # configs[1]: the headline window
    "window8x2000": dict(n_frames=8, points_per_frame=2000),
    # configs[3]: large window used for the HBM roofline and multi-GPU scaling
    "window8x20000": dict(n_frames=8, points_per_frame=20000),
    # configs[4]: marginalisation
    "marginalize8x2000": dict(n_frames=8, points_per_frame=2000, marginalize_first=True),
}


def make_config(name: str, seed: int = 0, **over) -> SynthWindow:
    kw = dict(CONFIGS[name])
    kw.update(over)
    return make_window(seed=seed, **kw)
