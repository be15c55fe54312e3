import numpy as np, sys, ctypes as C
sys.path.insert(0, '/root/repo')
from dsopp_b200 import capi, synth
win = synth.make_window(n_frames=8, points_per_frame=2000, seed=0, ab_scale=0.0)
h = capi.upload_window(win)
for i in range(3):
    h.first_estimate()
    h.solve_lm(20.0, max_it=7, min_it=7, ftol=0.0, ptol=0.0)
out = (C.c_longlong * 16)()
capi.load_library().dpba_debug_lm_clocks(out)
v = np.array(out[:8])
print("phase cycles:", np.diff(v))
