"""Golden vectors from the REFERENCE'S OWN bundle adjustment (oracle/build_ref_pba.py) for every run in
tests/ref_pba_cases.RUNS -> tests/golden/ref_pba.npz.  Run in the build container (needs /root/reference); the GPU box and
any checkout without the reference use the committed file.

    python tools/make_ref_pba_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_pba_cases as RC  # noqa: E402
from oracle import build_ref_pba  # noqa: E402

# The golden file keeps a cross-section of what the sequences leave behind (every status / residual / weight / energy of the
# linearisation sweep, the Jacobian blocks of four pairs, every linear system, the solve and marginalisation results); the
# live comparison in tests/test_reference_pba.py covers every array whenever the library is present.
KEEP_JAC_PAIRS = ("res01", "res21")
BIG = ("du_t", "dv_t", "J_ref", "J_tgt", "du_id", "dv_id", "d_idepth", "Hpd")


def keep(key):
    tag, mid, leaf = (key.split("/") + [""])[:3] if key.count("/") == 2 else (key.split("/")[0], "", key.split("/")[-1])
    if mid == "":  # systems, energies, results
        return True
    if tag == "finish":
        return leaf in ("status", "cand", "e", "flags", "n_inliers", "rel_baseline", "T_lin", "ab0", "state_eps", "idepth")
    if tag in ("solve", "marg"):
        return leaf not in ("corrected", "Hpd", "bcs", "r", "jac_valid") or (leaf == "Hpd" and mid == "lm1")
    if tag == "lin" and mid.startswith("res"):
        if leaf in BIG:
            return mid in KEEP_JAC_PAIRS and leaf in ("J_ref", "J_tgt", "d_idepth")
        return leaf in ("status", "cand", "jac_valid", "r", "w", "e")
    if tag == "fej" and mid.startswith("res"):
        return mid in KEEP_JAC_PAIRS[:2] and leaf in ("du_t", "dv_t", "du_id", "dv_id", "jac_valid", "bcs")
    if tag == "fej" and mid.startswith("lm"):
        return leaf == "corrected" and mid == "lm1"
    if tag == "schur" and mid.startswith("lm"):
        return leaf in ("inv_hdd", "b_d", "ill") or (leaf == "Hpd" and mid == "lm1")
    if tag == "idepths" and mid.startswith("lm"):
        return leaf == "idepth_step"
    if tag == "nohuber" and mid.startswith("res"):
        return leaf in ("w", "e") and mid in KEEP_JAC_PAIRS
    return False


def main():
    assert build_ref_pba.have_reference(), "needs /root/reference"
    out = {}
    for name, (case, seq, kw) in RC.RUNS.items():
        ref = RC.run(RC.ReferenceBackend, case, seq, **kw)
        kept = {k: v for k, v in ref.items() if keep(k)}
        for k, v in kept.items():
            out[f"{name}::{k}"] = v
        print(f"{name}: {len(kept)} of {len(ref)} arrays kept")
    for name in RC.PA_CASES:
        ref = RC.pa_run_reference(name)
        ref["pa/uv"] = ref["pa/uv"].astype(np.int16)  # integer pixel positions
        ref["pa/idepth"] = ref["pa/idepth"][::7]  # a cross-section; the live test compares every landmark
        ref["pa/patch"] = ref["pa/patch"][::7]
        for k, v in ref.items():
            out[f"{name}::{k}"] = v
        print(f"{name}: {len(ref['pa/uv'])} landmarks, energy {ref['pa/result'][0]:.6f}")
    path = os.path.join(ROOT, "tests", "golden", "ref_pba.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
