"""Golden vectors from the REFERENCE'S OWN tracker arithmetic (oracle/build_ref_tracker.py: create_depth_maps.cpp whole,
landmarks_activator.cpp:122-316) on the windows of tests/ref_tracker_cases.py -> tests/golden/ref_tracker.npz.
Run in the build container (needs /root/reference); the GPU box uses the committed file.

    python tools/make_ref_tracker_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_tracker_cases as TC  # noqa: E402
from oracle import build_ref_tracker, ref_tracker  # noqa: E402


def main():
    assert build_ref_tracker.have_reference(), "needs /root/reference"
    out = {}
    win, frames, variances = TC.depth_case()
    tgt = frames[-1]
    poses = [f.t_world_agent() for f in frames]
    for tag, var in (("const", None), ("var", variances)):
        maps = ref_tracker.create_reference_depth_maps(poses, tgt.intr, tgt.W, tgt.H, TC.DEPTH_LEVELS,
                                                       TC.track_landmarks(frames, var))
        for l, (idw, wgt) in enumerate(maps):
            out[f"depth/{tag}/idepth{l}"], out[f"depth/{tag}/weight{l}"] = idw, wgt
        print("depth", tag, [int((w > 0).sum()) for _, w in maps])
    win, _, images, masks, cands = TC.activation_case()
    T = [f.T_w_lin for f in win.frames]
    e, ab = [f.exposure for f in win.frames], [f.ab0 for f in win.frames]
    status, idepth = [], []
    for r, l, rho0, min_inl, sigma in cands:
        f = win.frames[r]
        s, rho = ref_tracker.optimize_immature_landmark(T, e, ab, images, masks, f.intr, r, f.uv[l], f.patch[l], rho0, rho0,
                                                        min_inl, sigma)
        status.append(s), idepth.append(rho)
    out["activation/status"], out["activation/idepth"] = np.array(status, np.uint8), np.array(idepth)
    print("activation", np.bincount(out["activation/status"], minlength=3))
    path = os.path.join(ROOT, "tests", "golden", "ref_tracker.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
