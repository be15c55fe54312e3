"""Where the end-to-end step of bench.py spends its time (run under gpurun): wall-clock per phase with a device
synchronisation after each phase, plus the raw pinned H2D bandwidth of this box for scale."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsopp_b200 import capi, synth
import bench

win = bench.build_window(1)
n = win.n_frames
h = capi.upload_window(win)
eps0 = np.concatenate([f.state_eps for f in win.frames])
keep, host_frames, host_status = [], [], []
for i, f in enumerate(win.frames):
    arrs = [bench.pin(a) for a in (f.image.astype(np.float32), f.mask, f.uv.astype(np.float32), f.idepth.astype(np.float32), f.patch.astype(np.float32), f.flags)]
    keep += [a[1] for a in arrs]
    host_frames.append((f,) + tuple(a[0] for a in arrs))
    rows = {}
    for t in range(n):
        if t != i:
            a, tk = bench.pin(win.statuses[(i, t)]); keep.append(tk); rows[t] = a
    host_status.append(rows)

def sync():
    torch.cuda.synchronize()

def step(tm):
    t0 = time.perf_counter()
    for _ in range(h.n_frames):
        h.remove_frame(0)
    for (f, img, msk, uv, idp, pat, flg) in host_frames:
        h.push_frame(f.frame_id, img, msk, f.T_w_lin, f.exposure, f.ab0, f.intr, f.fixed)
    t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
    for i, (f, img, msk, uv, idp, pat, flg) in enumerate(host_frames):
        h.set_landmarks(i, uv, idp, pat, flg)
        h.set_frame_statuses(i, host_status[i])
    h.set_state(eps0, np.zeros_like(eps0))
    t3 = time.perf_counter(); sync(); t4 = time.perf_counter()
    h.first_estimate()
    h.solve_lm(20.0, max_it=7, min_it=7, ftol=0.0, ptol=0.0)
    t5 = time.perf_counter()
    h.get_state()
    for i in range(n):
        h.get_landmarks(i)
        h.get_frame_statuses(i)
    t6 = time.perf_counter()
    tm.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t6 - t0])

tm = []
for i in range(13):
    step(tm)
tm = np.array(tm[3:]) * 1e3
print("ms: push_frame calls %.3f | drain %.3f | landmarks+statuses calls %.3f | drain %.3f | first_estimate+solve %.3f | readback %.3f | total %.3f" % tuple(np.median(tm, axis=0)))
# raw H2D for scale
src = torch.empty(32 << 20, dtype=torch.uint8).pin_memory(); dst = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): dst.copy_(src, non_blocking=True)
sync(); t0 = time.perf_counter()
for _ in range(10): dst.copy_(src, non_blocking=True)
sync(); dt = (time.perf_counter() - t0) / 10
print("raw pinned H2D 32 MiB: %.3f ms = %.1f GB/s" % (dt * 1e3, (32 << 20) / dt / 1e9))
