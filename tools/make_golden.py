#!/usr/bin/env python
"""Generates tests/golden/pba_window_3x48.npz: a small seeded window (inputs) and what the NumPy float64 oracle
computes on it (outputs).  The reference ships no golden vectors for this path and cannot be
built here, so these are OUR oracle's numbers, frozen: they pin the oracle against drift and give the GPU parity
tests constants to compare with.  Re-run only when the oracle is deliberately changed:

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dsopp_b200 import synth  # noqa: E402
from oracle import pba_oracle as O  # noqa: E402

SIGMA = 20.0
AB_REG = (10.0, 1e-2)  # weak affine prior so that the energy trace is dominated by the photometric term
OUT = os.path.join(ROOT, "tests", "golden", "pba_window_3x48.npz")


def main():
    win = synth.make_window(n_frames=3, points_per_frame=48, width=160, height=120, seed=2026, ab_scale=1.0)
    # a few non-trivial bookkeeping cases: a non-kOk residual, a marginalised landmark, an out-of-range idepth
    win.statuses[(0, 1)][3] = O.K_OUTLIER
    win.statuses[(2, 0)][5] = O.K_OOB
    win.frames[1].flags[7] = synth.FLAG_MARGINALIZED
    win.frames[2].idepth[9] = -1.0
    d = {"ab_reg": np.array(AB_REG), "sigma": SIGMA, "n_frames": win.n_frames, "width": win.width, "height": win.height}
    for i, f in enumerate(win.frames):
        d[f"f{i}_id"] = f.frame_id
        d[f"f{i}_timestamp"] = f.timestamp
        d[f"f{i}_T_w_lin"] = f.T_w_lin
        d[f"f{i}_exposure"] = f.exposure
        d[f"f{i}_ab0"] = f.ab0
        d[f"f{i}_intr"] = f.intr
        d[f"f{i}_image"] = f.image.astype(np.float32)
        d[f"f{i}_mask"] = f.mask
        d[f"f{i}_fixed"] = f.fixed
        d[f"f{i}_state_eps"] = f.state_eps
        d[f"f{i}_uv"] = f.uv
        d[f"f{i}_idepth"] = f.idepth
        d[f"f{i}_patch"] = f.patch
        d[f"f{i}_flags"] = f.flags
    for (r, t), st in win.statuses.items():
        d[f"status_{r}_{t}"] = st

    frames = O.frames_from_window(win)
    O.first_estimate_jacobians(frames)
    O.evaluate_jacobians(frames, SIGMA, fej=True, evaluate_jacobians=True, new_point=True, huber=True)
    for r, ref in enumerate(frames):
        for t, tgt in enumerate(frames):
            if r == t:
                continue
            res = ref.residuals[tgt.id]
            for k in ("r", "J_ref", "J_tgt", "d_idepth", "w", "e", "cand", "jac_valid"):
                d[f"out_{r}_{t}_{k}"] = getattr(res, k)
    Hp, bp = O.pose_pose(frames)
    Hs, bs = O.schur_complement(frames)
    d.update(out_Hp=Hp, out_bp=bp, out_Hs=Hs, out_bs=bs)
    for i, f in enumerate(frames):
        d[f"out_f{i}_Hpd"], d[f"out_f{i}_b_d"], d[f"out_f{i}_inv_hdd"], d[f"out_f{i}_ill"] = f.Hpd, f.b_d, f.inv_hdd, f.ill
    e, n = O.landmarks_energy(frames)
    d.update(out_energy=e, out_n_valid=n)
    # fixed-work LM solve (7 forced iterations, tolerances 0)
    frames = O.frames_from_window(win)
    O.first_estimate_jacobians(frames)
    trace = []
    e_lm, n_lm, _ = O.lm_solve(O.Problem(frames, SIGMA, ab_reg=AB_REG), O.LMOptions(7, 1e-5, 0.0, 0.0, True, 7, 1.0, 1.0), trace)
    d.update(out_lm_energy=e_lm, out_lm_n=n_lm, out_lm_trace=np.array([t["energy"] for t in trace]),
             out_lm_state=O.state_eps_stacked(frames))
    for i, f in enumerate(frames):
        d[f"out_lm_f{i}_idepth"] = f.idepth
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **d)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1024:.0f} KiB, {len(d)} arrays; LM energies {d['out_lm_trace']}")


if __name__ == "__main__":
    main()
