"""Where the end-to-end step (dpbah_solve_window, bench.py's `e2e`) spends its time: host wall-clock per phase, once with
the asynchronous uploads left in flight (what bench.py times) and once with the stream drained after every phase.
   python tools/e2e_breakdown.py [--raw | --records]   (default: intensity planes, bench.py's `e2e`)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi, host, synth  # noqa: E402

PHASES = ["remove", "push_frames", "landmarks+statuses+state", "first_estimate+solve", "readback"]


def main():
    raw = "--raw" in sys.argv
    records = "--records" in sys.argv
    win = synth.make_window(n_frames=8, points_per_frame=2000, seed=0, ab_scale=0.0)
    h = capi.upload_window(win)
    keep = []

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(t)
        return t.numpy()

    n = win.n_frames
    frames = [dict(frame_id=f.frame_id, image=pinned(f.image.astype(np.float32) if records else f.image[..., 0].astype(np.float32)),
                   mask=pinned(f.mask), T_w_lin=f.T_w_lin,
                   exposure=f.exposure, ab0=f.ab0, intr=f.intr, fixed=f.fixed, uv=pinned(f.uv.astype(np.float32)),
                   idepth=pinned(f.idepth.astype(np.float32)), patch=pinned(f.patch.astype(np.float32)), flags=pinned(f.flags))
              for f in win.frames]
    st = {k: pinned(v) for k, v in win.statuses.items()}
    eps0 = np.concatenate([f.state_eps for f in win.frames])
    kw = dict(max_it=7, min_it=7, ftol=0.0, ptol=0.0, alloc=lambda s, d: pinned(np.zeros(s, d)))
    if raw:
        kw.update(raw_gray=[pinned(np.clip(np.rint(f.image[..., 0]), 0, 255).astype(np.uint8)) for f in win.frames],
                  photometric_lut=np.arange(256, dtype=np.float32))
    io = host.WindowStep(h, frames, st, eps0, **kw)
    for sync in (0, 1):
        io.io.sync_phases = sync
        acc = np.zeros(5)
        for i in range(3 + 20):
            io.run()
            if i >= 3:
                acc += np.array(list(io.io.phase_ms))
        acc /= 20
        print(f"{'raw 8-bit frames' if raw else ('{I,dx,dy} records' if records else 'intensity planes')}, sync after every phase = {sync}: total {acc.sum():.3f} ms; " +
              ", ".join(f"{p} {v:.3f}" for p, v in zip(PHASES, acc)))
    print("h2d bytes", io.io.h2d_bytes, "d2h bytes", io.io.d2h_bytes)
    h.close()


if __name__ == "__main__":
    main()
