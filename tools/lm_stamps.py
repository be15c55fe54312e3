"""clock64() timeline of the single-CTA LM kernel (k_lm_solve) on the bench window: python tools/lm_stamps.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi, synth
win = synth.make_window(n_frames=8, points_per_frame=2000, seed=0, ab_scale=0.0)
h = capi.upload_window(win)
lib = capi.load_library()
out = np.zeros(64, np.int64)
lib.dpba_debug_stamps(1, None)
for rep in range(3):
    h.first_estimate()
    h.solve_lm(20.0, max_it=7, min_it=7, ftol=0.0, ptol=0.0)
lib.dpba_debug_stamps(1, out.ctypes.data)
t = out - out[0]
names = {0: "start", 1: "energy decision done", 2: "system filled"}
names.update({3 + k: f"block step {k} begins" for k in range(8)})
names.update({18: "factorised", 20: "step written", 21: "pair constants written"})
print("k_lm_solve energy body (last launch): sums", out[51] - out[50], "priors+decision", out[52] - out[51], " kernel start->body", out[50] - out[0])
print("k_reduce_system assemble block 0: core reduction", out[31] - out[30], "products", out[32] - out[31], "barrier", out[33] - out[32], "write", out[34] - out[33])
print("k_reduce_system schur block: sums", out[41] - out[40], "barrier", out[42] - out[41], "final", out[43] - out[42])
print("block step 0: loads + barrier + diagonal block + row recurrence + stores", out[13] - out[3], "barrier", out[14] - out[13], "trailing update", out[4] - out[14])
prev = 0
for i in sorted(names):
    print(f"{names[i]:28s} {t[i]:8d} cycles  (+{t[i] - prev})")
    prev = t[i]
h.close()
