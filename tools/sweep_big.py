"""The fused linearise at configs[3]'s 1-GPU shape (8 keyframes x 20000 points, 1.12 M patch-residuals), a few launches in
a row: the target of the ncu capture `-k regex:k_linearize_fused2 -s 2 -c 1` (tools/profile.sh) and a quick timing."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi, synth
win = synth.make_window(n_frames=8, points_per_frame=20000, seed=1, ab_scale=0.0)
h = capi.upload_window(win)
for kv in sys.argv[1:]:  # e.g. fused_epilogue=0 fused_lpb_max=128
    k, v = kv.split("=")
    h.set_option(k, int(v))
h.first_estimate()
h.profile_enable(True)
for i in range(6):
    if i == 2:
        h.profile_enable(True)
    h.linearize(20.0, True, True, False)
ms, n = h.profile_read()["linearize_fused"]
units = win.units
bytes_alg = 51 * units + 12 * 8 * 640 * 480 + 2 * (64 * 64 + 64) * 8
print(" ".join(sys.argv[1:]) or "defaults", end=": ")
print(f"k_linearize_fused2 at {units} patch-residuals: {1e3 * ms / n:.1f} us per launch, {units / (ms / n * 1e-3) / 1e9:.2f} G patch-residuals/s, "
      f"{bytes_alg / (ms / n * 1e-3) / 1e9:.0f} GB/s algorithmic")
h.close()
