"""Shared by tests/test_reference_pba.py and tools/make_ref_pba_golden.py: the windows and the call sequences on which
oracle/pba_oracle.py is pinned against the reference's own code (oracle/build_ref_pba.py).

A *case* is a seeded synthetic window plus edits that provoke the branches the reference has (out-of-bounds targets, masked
pixels, invalid inverse depths, non-OK connection statuses, marginalised / to-be-marginalised / outlier landmarks, pending
steps, Huber weights below one).  A *sequence* runs the same reference-named operations on either backend and collects
everything they leave behind, so the two result dictionaries can be compared key by key.
"""
import numpy as np

from dsopp_b200 import synth
from oracle import features_oracle as FO
from oracle import pba_oracle as O

AB_REG = (1e12, 1e8)  # fabric.cpp:68-69
FIXED_REG = 1e12
SIGMA = 20.0


# ----------------------------------------------------------------------------------------------------------------------
# cases
# ----------------------------------------------------------------------------------------------------------------------
def _raws(win):
    # the synthetic generator stores float32 {I, dx, dy}; the reference's PixelMap derives dx, dy itself in Precision
    return [np.ascontiguousarray(f.image[..., 0], dtype=np.float64) for f in win.frames]


def case_plain(seed=3):
    win = synth.make_window(n_frames=4, points_per_frame=80, width=160, height=120, seed=seed)
    return win, _raws(win), {}


def case_edge(seed=11):
    """Every branch of evaluateJacobians / firstEstimateJacobians_ / the block evaluation at least a few times."""
    win = synth.make_window(n_frames=5, points_per_frame=90, width=160, height=120, seed=seed, pose_noise=4e-3,
                            idepth_noise=5e-3)
    rng = np.random.default_rng(seed + 1000)
    H, W = win.height, win.width
    steps = {}
    for k, f in enumerate(win.frames):
        M = len(f.idepth)
        # landmarks whose own pattern leaves the ROI, and others whose reprojection does
        f.uv[0] = [5.0, 60.0]
        f.uv[1] = [W - 6.0, 40.0]
        f.uv[2] = [80.0, 5.0]
        f.uv[3] = [70.0, H - 6.0]
        f.uv[4] = [6.0, 6.0]  # exactly on the border: inside
        f.uv[5] = [W - 7.0, H - 7.0]
        # invalid / extreme inverse depths (camera_model_base.hpp:52-58)
        f.idepth[6] = -2e-4
        f.idepth[7] = -0.5e-4
        f.idepth[8] = 1011.0
        f.idepth[9] = 1e-7
        f.idepth[10] = 3.0  # large parallax: most targets out of bounds
        # landmark flags: marginalised, to-be-marginalised, outlier
        f.flags[11:14] = synth.FLAG_MARGINALIZED
        f.flags[14:17] = synth.FLAG_MARGINALIZED | synth.FLAG_TO_MARGINALIZE
        f.flags[17:19] = 4
        if k in (1, 3):
            f.mask[30:70, 50:110] = 0  # static mask on two targets
            f.mask[::7, ::5] = 0
        steps[k] = (rng.uniform(-1, 1, 8) * np.array([2e-3] * 6 + [1e-3, 0.2]) * (k > 0), rng.uniform(-1, 1, M) * 2e-3)
    for (r, t), st in win.statuses.items():
        M = len(st)
        st[19:22] = O.K_OUTLIER
        st[22:24] = O.K_OOB
        st[24:26] = O.K_UNKNOWN
        st[26] = O.K_OCCLUDED
        st[rng.integers(27, M, 4)] = rng.integers(1, 5, 4)
    return win, _raws(win), dict(steps=steps)


def case_edge0(seed=11):
    """case_edge without the pending steps (the C ABI has no setter for a pending inverse-depth step, so the CUDA path is
    compared with the reference on this variant, tests/test_gpu_reference_golden.py)."""
    win, raws, _ = case_edge(seed)
    return win, raws, {}


def case_marginalise(seed=5):
    """BASELINE configs[4] in small: every landmark of keyframe 0 flagged, keyframe 0 itself leaves the window."""
    win = synth.make_window(n_frames=4, points_per_frame=70, width=160, height=120, seed=seed, marginalize_first=True)
    return win, _raws(win), dict(frame_to_marginalize=0)


CASES = {"plain": case_plain, "edge": case_edge, "edge0": case_edge0, "marginalise": case_marginalise}


# ----------------------------------------------------------------------------------------------------------------------
# the two backends behind one vocabulary
# ----------------------------------------------------------------------------------------------------------------------
class OracleBackend:
    name = "oracle"

    def __init__(self, win, raws, extra):
        self.frames = []
        for f, raw in zip(win.frames, raws):
            fr = O.Frame(f.frame_id, f.timestamp, f.T_w_lin, f.exposure, f.ab0, f.intr, FO.pixel_info(raw), f.mask, f.fixed,
                         f.uv, f.idepth, f.patch, f.flags, f.state_eps)
            self.frames.append(fr)
        for (r, t), st in win.statuses.items():
            self.frames[r].residuals[self.frames[t].id] = O.Residuals(st)
        for k, (ds, di) in extra.get("steps", {}).items():
            self.frames[k].state_eps_step = np.array(ds)
            self.frames[k].idepth_step = np.array(di)
        if "frame_to_marginalize" in extra:
            self.frames[extra["frame_to_marginalize"]].to_marginalize = True
        n = O.BLOCK * len(self.frames)
        self.H_marg, self.b_marg, self.e_marg = np.zeros((n, n)), np.zeros(n), 0.0

    @property
    def n(self):
        return len(self.frames)

    def first_estimate(self):
        O.first_estimate_jacobians(self.frames)

    def evaluate(self, fej, jac, huber, sigma):
        O.evaluate_jacobians(self.frames, sigma if huber else 0.0, fej=fej, evaluate_jacobians=jac, huber=huber)

    def change_statuses(self, accept=True):
        O.change_residual_statuses(self.frames, accept)

    def residuals(self, r, t):
        x = self.frames[r].residuals[self.frames[t].id]
        return dict(status=x.status.copy(), cand=x.cand.copy(), r=x.r.copy(), du_id=x.du_id.copy(), dv_id=x.dv_id.copy(),
                    du_t=x.du_t.copy(), dv_t=x.dv_t.copy(), jac_valid=x.jac_valid.copy(), J_ref=x.J_ref.copy(),
                    J_tgt=x.J_tgt.copy(), d_idepth=x.d_idepth.copy(), w=x.w.copy(), e=x.e.copy(), bcs=x.bcs.copy())

    def landmarks(self, f):
        fr = self.frames[f]
        D = O.BLOCK * self.n
        Hpd = fr.Hpd if fr.Hpd.shape[1] == D else np.zeros((len(fr.idepth), D))
        return dict(idepth=fr.idepth.copy(), idepth_step=fr.idepth_step.copy(), inv_hdd=fr.inv_hdd.copy(), b_d=fr.b_d.copy(),
                    Hpd=np.array(Hpd), ill=fr.ill.copy(), corrected=fr.corrected.copy(), ref_pattern=fr.ref_pattern.copy(),
                    flags=(fr.lm_marginalized * 1 + fr.lm_to_marginalize * 2 + fr.lm_outlier * 4).astype(np.uint8))

    def frame_state(self, f):
        fr = self.frames[f]
        return fr.state_eps.copy(), fr.state_eps_step.copy(), fr.t_world_agent()[:3, :4]

    def linear_systems(self, for_marginalized=False, prior=None):
        Hp, bp = O.pose_pose(self.frames, for_marginalized)
        Hs, bs = O.schur_complement(self.frames, for_marginalized)
        if prior is not None:
            O.linear_system_prior(self.frames, Hp, bp, np.asarray(prior[0], float), prior[1], for_marginalized)
        return Hp, bp, Hs, bs

    def calculate_idepths(self, step, lam):
        O.calculate_idepths(self.frames, np.asarray(step, float), lam)

    def landmarks_energy(self, for_marginalized=False):
        e, n = O.landmarks_energy(self.frames, for_marginalized)
        return float(e), int(n)

    def solve(self, fej, max_iterations, force_accept, sigma):
        opt = O.LMOptions(max_iterations, 1.0 / 1e5, 1e-8, 1e-8, force_accept, 3, 1.0, 1.0)
        prob = O.Problem(self.frames, sigma, self.H_marg, self.b_marg, self.e_marg, AB_REG, FIXED_REG, fej)
        if fej:
            O.first_estimate_jacobians(self.frames)
        e, n, c = O.lm_solve(prob, opt)
        return float(e), int(n), bool(c)

    def finish_solve(self, sigma):
        """What EigenPhotometricBundleAdjustment::solve does after the LM loop without uncertainties
        (eigen_photometric_bundle_adjustment.cpp:88,99): relinearizeSystem, updatePointStatuses(1, sigma)."""
        last = self.frames[-1]  # photometric_bundle_adjustment.cpp:311-316
        last.T_lin = last.t_world_agent()
        last.ab0 = last.affine_brightness()
        last.state_eps = np.zeros(O.BLOCK)
        O.update_point_statuses(self.frames, 1, sigma)

    def extras(self, f):
        fr = self.frames[f]
        return dict(rel_baseline=fr.rel_baseline.copy(), n_inliers=fr.n_inliers.copy(), T_lin=fr.T_lin[:3, :4].copy(), ab0=fr.ab0.copy())

    def marginalize(self, fej, sigma):
        if fej:
            O.first_estimate_jacobians(self.frames)
        O.evaluate_jacobians(self.frames, sigma, fej=fej, evaluate_jacobians=True, huber=True)
        O.change_residual_statuses(self.frames)
        self.frames, self.H_marg, self.b_marg, self.e_marg = O.update_marginalized_linear_system(
            self.frames, self.H_marg, self.b_marg, self.e_marg, AB_REG, FIXED_REG)
        return self.H_marg, self.b_marg, float(self.e_marg)


class ReferenceBackend:
    name = "reference"

    def __init__(self, win, raws, extra):
        from oracle import ref_pba
        self.rw = ref_pba.window_from_synth(win, raws)
        for k, (ds, di) in extra.get("steps", {}).items():
            self.rw.set_frame_state(k, step=ds)
            self.rw.set_idepth_steps(k, di)
        if "frame_to_marginalize" in extra:
            self.rw.set_frame_flags(extra["frame_to_marginalize"], to_marginalize=True)
        self.idx = list(range(self.rw.n_frames))  # live position -> original frame index (for n_lm bookkeeping)

    @property
    def n(self):
        return self.rw.n_frames

    def first_estimate(self):
        self.rw.first_estimate()

    def evaluate(self, fej, jac, huber, sigma):
        self.rw.evaluate(fej, jac, huber, sigma)

    def change_statuses(self, accept=True):
        self.rw.change_statuses(accept)

    def residuals(self, r, t):
        return self.rw.residuals(r, t)

    def landmarks(self, f):
        return self.rw.landmarks(f)

    def frame_state(self, f):
        return self.rw.frame_state(f)

    def linear_systems(self, for_marginalized=False, prior=None):
        return self.rw.linear_systems(for_marginalized, prior)

    def calculate_idepths(self, step, lam):
        self.rw.calculate_idepths(step, lam)

    def landmarks_energy(self, for_marginalized=False):
        return self.rw.landmarks_energy(for_marginalized)

    def solve(self, fej, max_iterations, force_accept, sigma):
        return self.rw.solve(fej=fej, max_iterations=max_iterations, force_accept=force_accept, sigma_huber=sigma,
                             affine_reg=AB_REG, fixed_reg=FIXED_REG)

    def finish_solve(self, sigma):
        self.rw.relinearize_system()
        self.rw.update_point_statuses(1, sigma)

    def extras(self, f):
        T, ab = self.rw.linearization_point(f)
        return dict(rel_baseline=self.rw.relative_baseline(f), n_inliers=self.rw.landmarks(f)["n_inliers"], T_lin=T, ab0=ab)

    def marginalize(self, fej, sigma):
        before = self.rw.n_frames
        out = self.rw.marginalize(fej=fej, sigma_huber=sigma, affine_reg=AB_REG, fixed_reg=FIXED_REG)
        if self.rw.n_frames != before:  # frames flagged to_marginalize left the deque (problem.hpp:200-202)
            gone = before - self.rw.n_frames
            self.rw.n_lm = self.rw.n_lm[gone:]  # the cases only ever drop the oldest frames
        return out


# ----------------------------------------------------------------------------------------------------------------------
# sequences
# ----------------------------------------------------------------------------------------------------------------------
def _snapshot(b, out, tag, with_jac=True):
    n = b.n
    for r in range(n):
        for t in range(n):
            if r == t:
                continue
            res = b.residuals(r, t)
            keys = ["status", "cand", "jac_valid", "r", "w", "e", "bcs"]
            if with_jac:
                keys += ["du_id", "dv_id", "du_t", "dv_t", "J_ref", "J_tgt", "d_idepth"]
            for k in keys:
                out[f"{tag}/res{r}{t}/{k}"] = np.asarray(res[k])
    for f in range(n):
        lm = b.landmarks(f)
        for k in ("idepth", "idepth_step", "inv_hdd", "b_d", "Hpd", "ill", "corrected", "flags"):
            out[f"{tag}/lm{f}/{k}"] = np.asarray(lm[k])
        se, st, T = b.frame_state(f)
        out[f"{tag}/frame{f}/state_eps"], out[f"{tag}/frame{f}/step"], out[f"{tag}/frame{f}/T"] = se, st, np.asarray(T)


def seq_linearize(b, fej, sigma=SIGMA):
    """firstEstimateJacobians -> calculateEnergy's sweep -> linearize's sweep -> both systems (+ priors) -> idepth steps."""
    out = {}
    if fej:
        b.first_estimate()
        _snapshot(b, out, "fej")
    b.evaluate(fej, False, True, sigma)  # problem.hpp:290-292
    _snapshot(b, out, "energy", with_jac=False)
    out["energy/landmarks_energy"] = np.array(b.landmarks_energy())
    b.evaluate(fej, True, True, sigma)  # problem.hpp:322-324
    _snapshot(b, out, "lin")
    Hp, bp, _, _ = b.linear_systems()
    out["lin/H_pose_noprior"], out["lin/b_pose_noprior"] = Hp, bp
    Hp, bp, Hs, bs = b.linear_systems(prior=(AB_REG, FIXED_REG))
    out["lin/H_pose"], out["lin/b_pose"], out["lin/H_schur"], out["lin/b_schur"] = Hp, bp, Hs, bs
    _snapshot(b, out, "schur", with_jac=False)  # inv_hdd, b_d, Hpd, ill left behind by the Schur complement
    rng = np.random.default_rng(7)
    step = rng.uniform(-1, 1, len(bp)) * 1e-3
    b.calculate_idepths(step, 1e-5)
    _snapshot(b, out, "idepths", with_jac=False)
    b.change_statuses(True)
    b.evaluate(fej, True, False, 0.0)  # the no-Huber instantiation of covarianceMatrixPosePose (problem.hpp:232-233)
    _snapshot(b, out, "nohuber")
    Hp, bp, Hs, bs = b.linear_systems()
    out["nohuber/H_pose"], out["nohuber/b_pose"], out["nohuber/H_schur"], out["nohuber/b_schur"] = Hp, bp, Hs, bs
    return out


def seq_solve(b, fej, max_iterations=7, force_accept=True, sigma=SIGMA):
    """EigenPhotometricBundleAdjustment::solve up to the LM result (eigen_photometric_bundle_adjustment.cpp:56-84)."""
    out = {}
    e, n, c = b.solve(fej, max_iterations, force_accept, sigma)
    out["solve/result"] = np.array([e, n, float(c)])
    _snapshot(b, out, "solve", with_jac=False)
    # ... and what follows the loop: relinearizeSystem + updatePointStatuses (the 75 % energy quantile, outlier resets,
    # inlier counts, relative baselines)
    b.finish_solve(sigma)
    _snapshot(b, out, "finish", with_jac=False)
    for f in range(b.n):
        for k, v in b.extras(f).items():
            out[f"finish/lm{f}/{k}" if k in ("rel_baseline", "n_inliers") else f"finish/frame{f}/{k}"] = np.asarray(v)
    return out


def seq_marginalise(b, fej=True, sigma=SIGMA):
    """pushFrame's marginalisation update (eigen_photometric_bundle_adjustment.cpp:121-130), then a solve on the prior."""
    out = {}
    H, bb, e = b.marginalize(fej, sigma)
    out["marg/H"], out["marg/b"], out["marg/energy"] = np.array(H), np.array(bb), np.array([e])
    out["marg/n_frames"] = np.array([b.n])
    _snapshot(b, out, "marg", with_jac=False)
    out.update(seq_solve(b, fej, 4, True, sigma))
    return out


def run(backend_cls, case, sequence, **kw):
    win, raws, extra = CASES[case]()
    b = backend_cls(win, raws, extra)
    return {"linearize": seq_linearize, "solve": seq_solve, "marginalise": seq_marginalise}[sequence](b, **kw)


# (case, sequence, keyword arguments) -> the golden file holds the reference's output of each
RUNS = {
    "plain_lin_fej": ("plain", "linearize", dict(fej=True)),
    "plain_lin_nofej": ("plain", "linearize", dict(fej=False)),
    "edge_lin_fej": ("edge", "linearize", dict(fej=True)),
    "edge_lin_nofej": ("edge", "linearize", dict(fej=False, sigma=6.0)),
    "plain_solve_fej": ("plain", "solve", dict(fej=True)),
    "plain_solve_nofej_lm": ("plain", "solve", dict(fej=False, max_iterations=10, force_accept=False)),
    "edge_solve_fej": ("edge", "solve", dict(fej=True)),
    "edge0_lin_fej": ("edge0", "linearize", dict(fej=True)),
    "edge0_solve_fej": ("edge0", "solve", dict(fej=True)),
    "marginalise": ("marginalise", "marginalise", dict(fej=True)),
}

# integer / boolean keys are compared exactly, the rest relatively to the array's largest magnitude
EXACT = ("status", "cand", "jac_valid", "ill", "flags", "n_frames", "n_inliers")


def compare(ref, got, rtol=1e-9, skip=()):
    """-> list of (key, error) that violate the bar; keys present on one side only are violations."""
    bad = []
    for k in sorted(set(ref) | set(got)):
        if any(s in k for s in skip):
            continue
        if k not in ref or k not in got:
            bad.append((k, "missing"))
            continue
        a, g = np.asarray(ref[k]), np.asarray(got[k])
        if a.shape != g.shape:
            bad.append((k, f"shape {a.shape} vs {g.shape}"))
            continue
        if k.rsplit("/", 1)[-1] in EXACT:
            if not np.array_equal(a.astype(np.int64), g.astype(np.int64)):
                bad.append((k, int((a.astype(np.int64) != g.astype(np.int64)).sum())))
            continue
        scale = max(float(np.abs(a).max()) if a.size else 0.0, 1e-300)
        err = float(np.abs(a.astype(float) - g.astype(float)).max()) / scale if a.size else 0.0
        if not err <= rtol:
            bad.append((k, err))
    return bad


# ----------------------------------------------------------------------------------------------------------------------
# coarse-tracker image alignment: the reference's PoseAlignerProblem under its LM driver
# ----------------------------------------------------------------------------------------------------------------------
PA_CASES = {  # name -> (seed, density, width, height, ab_scale, affine regulariser, masked target?)
    "pa_sparse": (3, 0.05, 160, 120, 0.0, (1e12, 1e8), False),
    "pa_masked_affine": (6, 0.3, 160, 120, 1.0, (10.0, 1e-2), True),
    "pa_dense": (8, 1.0, 96, 64, 0.0, (1e12, 1e8), False),
}


def pa_case(name):
    seed, density, W, H, ab_scale, reg, masked = PA_CASES[name]
    case = synth.make_alignment_case(seed=seed, width=W, height=H, density=density, pose_noise=4e-3, ab_scale=ab_scale)
    r, t = case.reference, case.target
    rraw, traw = r.image[..., 0].astype(np.float64), t.image[..., 0].astype(np.float64)
    mask = t.mask.copy()
    if masked:
        mask[20:40, 30:90] = 0
        mask[::9, ::4] = 0
    w, s = case.weight.copy(), case.idepth_sum.copy()
    w[10, 10], s[10, 10] = 2.0, 1e-7  # inverse depth below kMinIdepth: no landmark (local_frame.hpp:379-381)
    w[2, 50] = 1.0  # inside the 4-px border: no landmark (:372-373)
    return dict(case=case, rraw=rraw, traw=traw, mask=mask, weight=w, idepth_sum=s, reg=reg)


def pa_run_oracle(name):
    from oracle import pose_alignment_oracle as PA
    c = pa_case(name)
    r, t = c["case"].reference, c["case"].target
    ref = PA.PAFrame(r.T_w_true, r.exposure, r.ab0, r.intr, FO.pixel_info(c["rraw"]), r.mask)
    tgt = PA.PAFrame(c["case"].T_w_target_guess, t.exposure, t.ab0, t.intr, FO.pixel_info(c["traw"]), c["mask"])
    uv, idepth, patch = PA.landmarks_from_depth_map(c["idepth_sum"], c["weight"], ref.image)
    a = PA.solve(ref, tgt, uv, idepth, patch, ab_reg=c["reg"])
    return {"pa/result": np.array([a["energy"], a["n_valid"], float(a["converged"])]), "pa/T_t_r": a["T_t_r"][:3, :4],
            "pa/ab_eps": a["ab_eps"], "pa/H": a["H"], "pa/uv": uv, "pa/idepth": idepth, "pa/patch": patch}


def pa_run_reference(name):
    from oracle import ref_pba
    c = pa_case(name)
    r, t = c["case"].reference, c["case"].target
    b = ref_pba.pose_alignment_solve(r.T_w_true, r.exposure, r.ab0, c["rraw"], c["case"].T_w_target_guess, t.exposure, t.ab0,
                                     c["traw"], r.intr, c["idepth_sum"], c["weight"], tgt_mask=c["mask"], affine_reg=c["reg"])
    return {"pa/result": np.array([b["energy"], b["n_valid"], float(b["converged"])]), "pa/T_t_r": b["T_t_r"][:3, :4],
            "pa/ab_eps": b["ab_eps"], "pa/H": b["H"], "pa/uv": b["uv"], "pa/idepth": b["idepth"], "pa/patch": b["patch"]}
