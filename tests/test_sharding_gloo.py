"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo.  Each rank linearises ITS landmark shard (the
NumPy oracle stands in for the kernels), the packed reduced system is summed with all_reduce exactly as
dpba_linearize does with NCCL, and the result must equal the unsharded window's system; landmark bookkeeping
(rank = l % G, local = l // G) must round-trip bit-exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dsopp_b200 import sharding, synth

SIGMA = 20.0


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, out):
    from oracle import pba_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    win = synth.make_window(n_frames=4, points_per_frame=101, seed=5)
    frames = O.frames_from_window(sharding.shard_window(win, rank, world))
    O.first_estimate_jacobians(frames)
    O.evaluate_jacobians(frames, SIGMA, fej=True, evaluate_jacobians=True, new_point=True, huber=True)
    Hp, bp = O.pose_pose(frames)
    Hs, bs = O.schur_complement(frames)
    e, n = O.landmarks_energy(frames)
    d = Hp.shape[0]
    packed = torch.from_numpy(np.concatenate([Hp.ravel(), bp, Hs.ravel(), bs, [e, float(n)]]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    # back-substitution is local: every rank uses the same (replicated) pose step
    step = np.linspace(-1e-3, 1e-3, d)
    O.calculate_idepths(frames, step, 1e-5)
    steps = [torch.zeros(len(sharding.shard_indices(101, r, world)), dtype=torch.float64) for r in range(world)]
    mine = torch.from_numpy(np.ascontiguousarray(frames[1].idepth_step))
    gathered = [torch.zeros(len(s), dtype=torch.float64) for s in steps]
    dist.all_gather(gathered, mine) if all(len(s) == len(mine) for s in steps) else None
    if rank == 0:
        np.save(out, packed.numpy())
    np.save(out + f".steps{rank}.npy", frames[1].idepth_step)
    dist.destroy_process_group()


def test_sharded_linearisation_sums_to_the_unsharded_system(tmp_path):
    from oracle import pba_oracle as O
    world, out = 2, str(tmp_path / "packed.npy")
    mp.spawn(worker, args=(world, free_port(), out), nprocs=world, join=True)
    packed = np.load(out)
    win = synth.make_window(n_frames=4, points_per_frame=101, seed=5)
    frames = O.frames_from_window(win)
    O.first_estimate_jacobians(frames)
    O.evaluate_jacobians(frames, SIGMA, fej=True, evaluate_jacobians=True, new_point=True, huber=True)
    Hp, bp = O.pose_pose(frames)
    Hs, bs = O.schur_complement(frames)
    e, n = O.landmarks_energy(frames)
    ref = np.concatenate([Hp.ravel(), bp, Hs.ravel(), bs, [e, float(n)]])
    assert np.allclose(packed, ref, rtol=1e-11, atol=1e-9 * np.abs(ref).max())
    assert packed[-1] == ref[-1]  # residual counts are exact
    # local back-substitution, gathered with the index map, equals the unsharded one
    step = np.linspace(-1e-3, 1e-3, Hp.shape[0])
    O.calculate_idepths(frames, step, 1e-5)
    parts = [np.load(out + f".steps{r}.npy") for r in range(world)]
    got = sharding.gather_landmark_array(parts, 101, world)
    assert np.allclose(got, frames[1].idepth_step, rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (101, 2), (2000, 4), (20000, 8), (7, 8)])
def test_index_bookkeeping_round_trips_exactly(n, world):
    seen = np.zeros(n, dtype=np.int64)
    for rank in range(world):
        g = sharding.shard_indices(n, rank, world)
        r, l = sharding.to_local(g, world)
        assert (r == rank).all() and (l == np.arange(len(g))).all()
        assert np.array_equal(sharding.to_global(r, l, world), g)
        seen[g] += 1
    assert (seen == 1).all()
    sizes = [len(sharding.shard_indices(n, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
