"""Where a sharded iteration's time goes: entry / exit of the LM-loop kernels of rank 0 (dpba_debug_kernel_times) on the
bench window, 2000 points per keyframe and GPU, NVLink mailbox exchange.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/lm_stamps_multi.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
win = synth.make_window(n_frames=8, points_per_frame=2000 * world, seed=0, ab_scale=0.0)
h = capi.upload_window(win, device=local, rank=rank, world_size=world)
capi.attach_peers(h, rank, world, dev)
h.set_option("peer_exchange", 1)
h.set_option("peer_fused", 0)
for name in sys.argv[1:]:
    k, v = name.split("=")
    h.set_option(k, int(v))
lib = capi.load_library()
h.set_option("debug_freeze_stamps", 6)  # stamps stop after the energy decision of the last full iteration
for rep in range(4):
    dist.barrier()
    lib.dpba_debug_stamps(1, None)
    h.first_estimate()
    h.solve_lm(20.0, max_it=7, min_it=7, ftol=0.0, ptol=0.0)
kt = np.zeros(32, np.int64)
lib.dpba_debug_kernel_times(kt.ctypes.data)
names = ["fused sweep", "core reduce", "energy decision", "schur reduce", "block assembly", "lm step", "back-substitution",
         "pair constants", "landmark accept", "scalar reduce", "mailbox exchange"]
rows = sorted((kt[2 * i], kt[2 * i + 1], names[i]) for i in range(len(names)) if kt[2 * i] > 0)
dist.barrier()
for r in range(world):
    if r == rank and r in (0, world - 1) and rows:
        t0 = rows[0][0]
        print(f"[rank {rank} of {world}] last launches of the LM-loop kernels, us from the first entry (entry -> exit):")
        for a, b, nm in rows:
            extra = f"   waiting for the peers {(kt[23] - kt[22]) / 1e3:5.1f} us" if nm == "mailbox exchange" else ""
            print(f"   {nm:18s} {(a - t0) / 1e3:8.1f} -> {(b - t0) / 1e3:8.1f}   ({(b - a) / 1e3:5.1f} us){extra}", flush=True)
    dist.barrier()
h.close()
dist.destroy_process_group()
