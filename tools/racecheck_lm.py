"""Smallest run that exercises every block step of the LM step's look-ahead LDL^T (8 keyframes -> 64 x 64 system), for
   compute-sanitizer --tool racecheck python tools/racecheck_lm.py
(no torch: the C ABI through ctypes only, so the sanitizer has little else to watch)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi, synth  # noqa: E402

win = synth.make_window(n_frames=8, points_per_frame=40, seed=3, width=160, height=120, ab_scale=0.0)
h = capi.upload_window(win)
h.first_estimate()
e, it, conv, nv = h.solve_lm(20.0, max_it=2, min_it=2, ftol=0.0, ptol=0.0)
print("energy", e, "iterations", it, "valid", nv)
eps, _ = h.get_state()
print("state checksum", float(np.abs(eps).sum()))
h.close()
