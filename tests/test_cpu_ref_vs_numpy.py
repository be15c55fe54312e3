"""The two independent CPU restatements (NumPy float64 oracle, C++ reference-dataflow port) must agree."""
import numpy as np
import pytest

from dsopp_b200 import synth
from oracle import cpu_ref
from oracle import pba_oracle as O

SIGMA = 20.0


def flagged(win):
    win.frames[0].flags[::5] = synth.FLAG_MARGINALIZED
    win.frames[0].flags[::10] = synth.FLAG_MARGINALIZED | synth.FLAG_TO_MARGINALIZE
    win.statuses[(1, 2)][::4] = O.K_OUTLIER
    win.frames[2].mask[200:300, 100:500] = 0


@pytest.mark.parametrize("fej", [True, False])
@pytest.mark.parametrize("for_marg", [False, True])
def test_double_port_matches_numpy_oracle(fej, for_marg):
    win = synth.make_window(n_frames=4, points_per_frame=120, seed=11)
    flagged(win)
    frames = O.frames_from_window(win)
    cw = cpu_ref.CpuWindow(win, use_float=False, threads=2)
    O.first_estimate_jacobians(frames)
    cw.first_estimate()
    O.evaluate_jacobians(frames, SIGMA, fej=fej, evaluate_jacobians=True, new_point=True, huber=True)
    cw.evaluate(SIGMA, fej, True)
    for r in range(4):
        for t in range(4):
            if r == t:
                continue
            res, got = frames[r].residuals[frames[t].id], cw.residuals(r, t)
            assert (got["cand"] == res.cand).all()
            for k, ref in (("r", res.r), ("J_ref", res.J_ref), ("J_tgt", res.J_tgt), ("d_idepth", res.d_idepth),
                           ("e", res.e), ("w", res.w)):
                assert np.allclose(got[k], ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max())), (r, t, k)
    Hp, bp = cw.pose_pose(for_marg)
    Hs, bs = cw.schur(for_marg)
    Hp_ref, bp_ref = O.pose_pose(frames, for_marg)
    Hs_ref, bs_ref = O.schur_complement(frames, for_marg)
    for a, b in ((Hp, Hp_ref), (bp, bp_ref), (Hs, Hs_ref), (bs, bs_ref)):
        assert np.abs(a - b).max() <= 1e-10 * max(np.abs(b).max(), 1e-30)
    if not for_marg:
        step = np.linalg.solve(Hp_ref + np.eye(32) * 1e6 - Hs_ref, bp_ref - bs_ref)
        O.calculate_idepths(frames, step, 1e-5)
        cw.calculate_idepths(step, 1e-5)
        for i, f in enumerate(frames):
            lm = cw.landmarks(i)
            assert np.allclose(lm["idepth_step"], f.idepth_step, rtol=1e-8, atol=1e-14)
            assert (lm["ill"].astype(bool) == f.ill).all()
    cw.close()


def test_float_port_statuses_match_double_away_from_borders():
    win = synth.make_window(n_frames=4, points_per_frame=400, seed=12)
    cd = cpu_ref.CpuWindow(win, use_float=False)
    cf = cpu_ref.CpuWindow(win, use_float=True)
    for c in (cd, cf):
        c.first_estimate()
        c.evaluate(SIGMA, True, True)
    flips = 0
    for r in range(4):
        for t in range(4):
            if r != t:
                a, b = cd.residuals(r, t), cf.residuals(r, t)
                flips += int((a["cand"] != b["cand"]).sum())
                same = a["cand"] == b["cand"]
                assert np.allclose(a["r"][same], b["r"][same], rtol=1e-3, atol=5e-2)
    assert flips <= 1
    cd.close(), cf.close()


def test_normal_solve_matches_numpy():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(40, 24))
    H = A.T @ A * 1e3
    H[:8, :8] += np.eye(8) * 1e16
    b = rng.normal(size=24) * 1e3
    x = cpu_ref.normal_solve(H, b)
    assert np.allclose(x, O.normal_solve(H, b), rtol=1e-8, atol=1e-12)


def test_gn_iteration_matches_numpy_lm_body():
    win = synth.make_window(n_frames=4, points_per_frame=150, seed=13, ab_scale=0.0)
    frames = O.frames_from_window(win)
    cw = cpu_ref.CpuWindow(win, use_float=False, threads=2)
    O.first_estimate_jacobians(frames)
    cw.first_estimate()
    prob = O.Problem(frames, SIGMA)
    for it in range(3):
        prob.linearize()
        step = prob.calculate_step(1e-5)
        e_ref, _ = prob.calculate_energy()
        prob.accept_step()
        e, times, st = cw.gn_iteration(SIGMA, True, 1e-5, (1e12, 1e8), 1e16)
        assert abs(e - e_ref) <= 1e-8 * abs(e_ref)
        assert np.linalg.norm(st - step) <= 1e-6 * np.linalg.norm(step) + 1e-12
        assert times[5] > 0
    cw.close()


def test_update_point_statuses_port_matches_numpy_oracle():
    """updatePointStatuses of the C++ restatement (double) against the NumPy oracle: threshold, reset residuals, inlier
    counts, relative baselines, outlier flags (photometric_bundle_adjustment.cpp:322-406)."""
    win = synth.make_window(n_frames=5, points_per_frame=200, seed=14)
    flagged(win)
    frames = O.frames_from_window(win)
    cw = cpu_ref.CpuWindow(win, use_float=False, threads=2)
    O.first_estimate_jacobians(frames)
    cw.first_estimate()
    O.evaluate_jacobians(frames, SIGMA, fej=True, evaluate_jacobians=False, new_point=True, huber=True)
    O.change_residual_statuses(frames)
    cw.evaluate(SIGMA, True, False)
    cw.change_statuses(True)
    thr_ref = O.update_point_statuses(frames, 1, SIGMA)
    thr = cw.update_point_statuses(1, SIGMA)
    assert abs(thr - thr_ref) <= 1e-9 * thr_ref
    for i, f in enumerate(frames):
        lf = cw.landmark_flags(i)
        act = ~f.lm_marginalized
        assert (lf["n_inliers"][act] == f.n_inliers[act]).all()
        assert (lf["outlier"].astype(bool)[act] == f.lm_outlier[act]).all()
        assert np.allclose(lf["rel_baseline"][act], f.rel_baseline[act], rtol=1e-9, atol=1e-12)
        for j, g in enumerate(frames):
            if i != j:
                got = cw.residuals(i, j)
                assert (got["status"] == f.residuals[g.id].status).all()
    cw.close()


def test_device_op_flavour_of_the_float_port_only_moves_last_bits():
    """`device_ops` (the kernels' operation order: tap chain of fused multiply-adds, 8-lane butterfly, correctly rounded
    Huber energy) must be the same arithmetic as the plain float build up to rounding, and leave the double build alone."""
    win = synth.make_window(n_frames=4, points_per_frame=300, seed=15)
    a = cpu_ref.CpuWindow(win, use_float=True)
    b = cpu_ref.CpuWindow(win, use_float=True)
    b.set_device_ops(True)
    d0 = cpu_ref.CpuWindow(win, use_float=False)
    d1 = cpu_ref.CpuWindow(win, use_float=False)
    d1.set_device_ops(True)
    for c in (a, b, d0, d1):
        c.first_estimate()
        c.evaluate(SIGMA, True, True)
    n_huber = 0
    for r in range(4):
        for t in range(4):
            if r != t:
                x, y = a.residuals(r, t), b.residuals(r, t)
                assert (x["cand"] == y["cand"]).all() and (a.jac_valid(r, t) == b.jac_valid(r, t)).all()
                assert np.allclose(x["r"], y["r"], rtol=0, atol=2e-4)      # intensities 0..255: a few ulp of 256
                assert np.allclose(x["e"], y["e"], rtol=2e-5, atol=1e-3)
                n_huber += int((y["w"] < 1).sum())
                for k in ("r", "e", "cand"):
                    assert np.array_equal(d0.residuals(r, t)[k], d1.residuals(r, t)[k])
    assert n_huber > 0  # the Huber branch was exercised
    for c in (a, b, d0, d1):
        c.close()
