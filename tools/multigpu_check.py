"""Run under torchrun on N GPUs: the sharded solve (landmarks dealt over the ranks, NCCL exchange of the packed
reduced system inside dpba_linearize / dpba_solve_lm) must reproduce the single-GPU solve of the same window.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

DPBA_SPEC_MULTI=1   also compares the one-allreduce speculative sequence with the two-sweep sequence
DPBA_PEER_EXCHANGE=1 also repeats everything with the NVLink mailbox all-reduce (peer_exchange.cu) instead of NCCL
(run under `timeout`: the kernel gives up after ~15 s per exchange if a peer never arrives, it does not hang)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi, sharding, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    def say(*a):
        print(f"[rank {rank}]", *a, flush=True)
    dist.init_process_group("nccl", device_id=dev)
    say("process group up")
    win = synth.make_window(n_frames=5, points_per_frame=601, seed=9, ab_scale=0.0)
    h = capi.upload_window(win, device=local, rank=rank, world_size=world)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    say("unique id broadcast")
    h.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    if os.environ.get("DPBA_SPEC_MULTI"):
        h.set_option("speculative_multi_gpu", int(os.environ["DPBA_SPEC_MULTI"]))
    say("dpba_comm_init done")
    h.first_estimate()
    Hp, bp, Hs, bs = h.linearize(20.0, True, True, False)
    say("linearize done")
    e, n = h.evaluate(20.0, True, True)
    say("evaluate done")
    h.first_estimate()
    E, it, conv, nv = h.solve_lm(20.0)
    say("solve_lm done")
    if os.environ.get("DPBA_SPEC_MULTI"):
        # the same solve through the other launch sequence (two sweeps + two allreduces per iteration) on the SAME
        # handle and communicator (an NCCL unique id serves one communicator): must agree
        eps_a, _ = h.get_state()
        shard = [sharding.shard_indices(len(f.idepth), rank, world) for f in win.frames]
        for i, f in enumerate(win.frames):
            h.set_landmarks(i, f.uv[shard[i]], f.idepth[shard[i]], f.patch[shard[i]], f.flags[shard[i]])
        for (r_, t_), st in win.statuses.items():
            h.set_statuses(r_, t_, st[shard[r_]])
        h.set_state(np.concatenate([f.state_eps for f in win.frames]), np.zeros(8 * win.n_frames))
        h.set_option("speculative_multi_gpu", 0)
        h.first_estimate()
        E2, it2, _, nv2 = h.solve_lm(20.0)
        eps2, _ = h.get_state()
        say(f"speculative multi-GPU E={E!r} it={it} nv={nv} | two-sweep multi-GPU E={E2!r} it={it2} nv={nv2} | max|d eps| {np.abs(eps_a - eps2).max():.2e}")
        assert abs(E - E2) <= 1e-6 * abs(E2) and it == it2 and nv == nv2 and np.abs(eps_a - eps2).max() < 1e-6
    if os.environ.get("DPBA_PEER_EXCHANGE"):
        # the same solve with the library's own NVLink mailbox all-reduce instead of ncclAllReduce, on the SAME handle:
        # every rank sums the same values in rank order, so state and energy must agree to rounding with the NCCL run
        eps_a, _ = h.get_state()
        capi.attach_peers(h, rank, world, dev)
        h.set_option("peer_exchange", 1)
        if os.environ.get("DPBA_PEER_FUSED") is not None:  # 0: stand-alone mailbox kernel (and with it the split exchange)
            h.set_option("peer_fused", int(os.environ["DPBA_PEER_FUSED"]))
        say("peers attached")
        shard = [sharding.shard_indices(len(f.idepth), rank, world) for f in win.frames]
        for i, f in enumerate(win.frames):
            h.set_landmarks(i, f.uv[shard[i]], f.idepth[shard[i]], f.patch[shard[i]], f.flags[shard[i]])
        for (r_, t_), st in win.statuses.items():
            h.set_statuses(r_, t_, st[shard[r_]])
        h.set_state(np.concatenate([f.state_eps for f in win.frames]), np.zeros(8 * win.n_frames))
        h.first_estimate()
        Hp3, bp3, Hs3, bs3 = h.linearize(20.0, True, True, False)
        for a_, b_, nm in ((Hp3, Hp, "Hp"), (bp3, bp, "bp"), (Hs3, Hs, "Hs"), (bs3, bs, "bs")):
            err = np.abs(a_ - b_).max() / max(np.abs(b_).max(), 1e-30)
            say(f"peer exchange {nm}: max|d|/max|nccl| = {err:.2e}")
            assert err < 1e-12
        e3, n3 = h.evaluate(20.0, True, True)
        assert abs(e3 - e) <= 1e-12 * abs(e) and n3 == n
        h.first_estimate()
        E3, it3, _, nv3 = h.solve_lm(20.0)
        eps3, _ = h.get_state()
        say(f"NCCL E={E!r} it={it} nv={nv} | peer exchange E={E3!r} it={it3} nv={nv3} | max|d eps| {np.abs(eps_a - eps3).max():.2e}")
        assert abs(E - E3) <= 1e-6 * abs(E) and it == it3 and nv == nv3 and np.abs(eps_a - eps3).max() < 1e-6
        # replicas must hold bitwise identical states (same sums in the same order on every rank)
        mine = torch.from_numpy(eps3.copy()).to(dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        assert all(torch.equal(every[0], t) for t in every), "replicated LM state differs between ranks"
    eps, _ = h.get_state()
    idepth = [h.get_landmarks(i)["idepth"] for i in range(win.n_frames)]
    ok = True

    def reset(hh, rk, ws):
        sh = [sharding.shard_indices(len(f.idepth), rk, ws) for f in win.frames]
        for i, f in enumerate(win.frames):
            hh.set_landmarks(i, f.uv[sh[i]], f.idepth[sh[i]], f.patch[sh[i]], f.flags[sh[i]])
        for (r_, t_), st in win.statuses.items():
            hh.set_statuses(r_, t_, st[sh[r_]])
        hh.set_state(np.concatenate([f.state_eps for f in win.frames]), np.zeros(8 * win.n_frames))
        hh.first_estimate()
        hh.evaluate(20.0, True, True)
        hh.change_residual_statuses(True)

    # updatePointStatuses over the union of the shards, from the initial window again: the device radix select sums its
    # 256-bin histograms over the ranks (identical inputs -> identical per-residual energies -> the very same threshold
    # float and inlier counts as the unsharded window)
    reset(h, rank, world)
    thr = h.update_point_statuses(1, 20.0)
    inl = [h.get_landmarks(i)["n_inliers"].copy() for i in range(win.n_frames)]
    say(f"update_point_statuses done, threshold {thr!r}")
    if rank == 0:
        ref = capi.upload_window(win, device=local)
        ref.first_estimate()
        Hp1, bp1, Hs1, bs1 = ref.linearize(20.0, True, True, False)
        e1, n1 = ref.evaluate(20.0, True, True)
        ref.first_estimate()
        E1, it1, _, nv1 = ref.solve_lm(20.0)
        eps1, _ = ref.get_state()
        for a, b, nm in ((Hp, Hp1, "Hp"), (bp, bp1, "bp"), (Hs, Hs1, "Hs"), (bs, bs1, "bs")):
            err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
            print(f"{nm}: max|d|/max|ref| = {err:.2e}")
            ok &= err < 1e-6
        print(f"energy {e} vs {e1}, n {n} vs {n1}; LM energy {E} vs {E1}, it {it} vs {it1}, max|d eps| {np.abs(eps - eps1).max():.2e}")
        ok &= abs(e - e1) <= 1e-6 * abs(e1) and n == n1
        ok &= abs(E - E1) <= 5e-5 * abs(E1) and abs(it - it1) <= 1 and np.abs(eps - eps1).max() < 2e-5
        for i in range(win.n_frames):
            full = ref.get_landmarks(i)["idepth"]
            mine = full[sharding.shard_indices(len(full), 0, world)]
            ok &= np.abs(mine - idepth[i]).max() < 5e-5
        reset(ref, 0, 1)
        thr1 = ref.update_point_statuses(1, 20.0)
        print(f"outlier threshold sharded {thr!r} vs unsharded {thr1!r}")
        ok &= bool(np.float32(thr) == np.float32(thr1))
        for i in range(win.n_frames):
            full = ref.get_landmarks(i)["n_inliers"]
            ok &= bool(np.array_equal(full[sharding.shard_indices(len(full), 0, world)], inl[i]))
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL", f"world={world}")
        ref.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    code = 0 if flag.item() else 1
    # ordered shutdown and a normal exit: the handle first (its captured graph, then its communicator), then torch's group
    h.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    say("done, exit code", code)
    sys.stdout.flush()
    sys.exit(code)


if __name__ == "__main__":
    main()
