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
print("k_reduce_system assemble block 0: core reduction", out[55] - out[54], "products", out[56] - out[55], "barrier", out[57] - out[56], "write", out[58] - out[57], "(merged_tail only)")
print("k_reduce_system schur block: sums", out[60] - out[59], "barrier", out[61] - out[60], "final", out[62] - out[61], "(merged_tail only)")
print("k_lm_step [cycles]: entry -> system filled", out[2] - out[22], "| LDL^T", out[18] - out[2], "| back substitution + step written", out[23] - out[18], "| total", out[23] - out[22])
print("block step 0: loads + barrier + diagonal block + row recurrence + stores", out[13] - out[3], "barrier", out[14] - out[13], "trailing update", out[4] - out[14])
f = out[30:42] - out[30]
print("k_linearize_fused2, CTA (0,0) thread 0 [cycles from kernel entry]: prologue barrier", f[1], "| group 0: pass 1 done", f[2], "pass 2 done", f[3], "warp reduce done", f[4],
      "| all groups done", f[5], "| barrier", f[6], "ref block + barrier", f[7], "finalise + barrier", f[8], "H_pd rows stored", f[9], "SYRK", f[10], "b vector, end", f[11])
prev = 0
for i in sorted(names):
    print(f"{names[i]:28s} {t[i]:8d} cycles  (+{t[i] - prev})")
    prev = t[i]

# per-CTA timeline of the last fused sweep (dpba_debug_cta_times): how long the CTAs take and how they share the SMs
ct = np.zeros(4096, np.int64)
lib.dpba_debug_cta_times(ct.ctypes.data, 4096)
ct = ct.reshape(1024, 4)
ct = ct[ct[:, 0] > 0]
if len(ct):
    t0 = ct[:, 0].min()
    ent, swp, end, sm = (ct[:, 0] - t0) / 1e3, (ct[:, 1] - ct[:, 0]) / 1e3, (ct[:, 2] - ct[:, 0]) / 1e3, ct[:, 3]
    per_sm = np.bincount(sm.astype(int))
    print(f"fused sweep, {len(ct)} CTAs: entry spread {ent.max():.1f} us, kernel span {(ct[:, 2].max() - t0) / 1e3:.1f} us; CTA time "
          f"min/median/max {end.min():.1f}/{np.median(end):.1f}/{end.max():.1f} us, sweep part {swp.min():.1f}/{np.median(swp):.1f}/{swp.max():.1f} us; "
          f"SMs with 1/2/3+ CTAs: {(per_sm == 1).sum()}/{(per_sm == 2).sum()}/{(per_sm > 2).sum()}")
    solo = np.isin(sm, np.nonzero(per_sm == 1)[0])
    if solo.any() and (~solo).any():
        print(f"   CTAs alone on their SM: median {np.median(end[solo]):.1f} us (sweep {np.median(swp[solo]):.1f}); sharing an SM: median {np.median(end[~solo]):.1f} us (sweep {np.median(swp[~solo]):.1f})")
    late = np.argsort(ct[:, 2])[-5:]
    print("   last five CTAs to finish: " + ", ".join(f"entry {ent[i]:.1f} + {end[i]:.1f} us (sweep {swp[i]:.1f}) on SM {int(sm[i])}" for i in late))

# entry / exit of the LM-loop kernels' last launches (dpba_debug_kernel_times): the gaps BETWEEN the kernels of an iteration;
# the stamps are frozen after the energy decision of iteration 6, so the table is one complete iteration (LM step -> energy)
h.set_option("debug_freeze_stamps", 6)
lib.dpba_debug_stamps(1, None)
h.first_estimate()
h.solve_lm(20.0, max_it=7, min_it=7, ftol=0.0, ptol=0.0)
kt = np.zeros(32, np.int64)
lib.dpba_debug_kernel_times(kt.ctypes.data)
knames = ["fused sweep", "core reduce", "energy decision", "schur reduce", "block assembly", "lm step", "back-substitution", "pair constants", "landmark accept"]
rows = sorted((kt[2 * i], kt[2 * i + 1], knames[i]) for i in range(len(knames)) if kt[2 * i] > 0)
if rows:
    t0 = rows[0][0]
    print("last launches of the LM-loop kernels, us from the first entry (entry -> exit):")
    for a, b, nm in rows:
        print(f"   {nm:18s} {(a - t0) / 1e3:8.1f} -> {(b - t0) / 1e3:8.1f}   ({(b - a) / 1e3:5.1f} us)")
h.close()
