"""GPU parity of the immature-landmark activation refine: dpba_refine_immature_landmarks against the
float64 oracle (oracle/activation_oracle.py) on the same candidates.  fp32 device arithmetic: refined inverse depths
95 % within 2e-4 relative, all within 1e-3 (1-D LM, 3 iterations), activate / delete decisions and valid-residual counts exact except for
candidates the oracle itself puts within fp32 noise of a decision boundary (counted, <= 1 %)."""
import numpy as np
import pytest

from dsopp_b200 import synth

pytestmark = pytest.mark.gpu


def test_refine_matches_oracle_and_improves_depths():
    from dsopp_b200 import capi
    from oracle import activation_oracle as A
    win = synth.make_window(n_frames=5, points_per_frame=150, seed=23, pose_noise=0.0, eps_scale=0.0, ab_scale=0.0)
    win.frames[2].mask[100:160, 200:330] = 0  # a masked region in one target
    h = capi.upload_window(win)
    fr = [A.ActFrame(f.frame_id, f.T_w_lin, f.exposure, f.ab0, f.intr, f.image, f.mask) for f in win.frames]
    rng = np.random.default_rng(1)
    total = flips = better = 0
    rel = []
    for r in (0, 3):
        f = win.frames[r]
        rho0 = (f.idepth_true * (1 + rng.uniform(-1, 1, len(f.idepth_true)) * 0.03)).astype(np.float32)
        rho0[:3] = [5.0, -0.5, 2000.0]  # hopeless / invalid candidates: must be deleted
        got_rho, got_act, got_n = h.refine_immature_landmarks(r, f.uv, rho0, f.patch, 3, 20.0)
        for l in range(len(rho0)):
            act, rho, n = A.optimize_immature_landmark(fr[r], fr, f.uv[l], f.patch[l], float(rho0[l]), 3, 20.0)
            total += 1
            if act != got_act[l] or n != got_n[l]:
                flips += 1
                continue
            if act:
                rel.append(abs(got_rho[l] - rho) / abs(rho))
                better += abs(got_rho[l] - f.idepth_true[l]) < abs(rho0[l] - f.idepth_true[l])
            else:
                assert got_rho[l] < 0 or got_n[l] < 3
        assert not got_act[:3].any()
    print(f"[activation] {total} candidates, {flips} decision flips, {better} improved")
    rel = np.array(rel)
    print(f"[activation] idepth rel err: median {np.median(rel):.1e}, 95% {np.quantile(rel, 0.95):.1e}, max {rel.max():.1e}")
    # 3 LM iterations in fp32: most candidates agree to ~1e-5; weakly observed depths (small H) amplify the fp32 residual
    # floor, as in the BA back-substitution
    assert np.quantile(rel, 0.95) <= 2e-4 and rel.max() <= 1e-3
    assert flips <= max(1, total // 100)
    assert better > 0.85 * (total - flips - 6)
    h.close()
