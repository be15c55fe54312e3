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


class ShardedProblem:
    """PhotometricBundleAdjustmentProblem over ONE rank's landmark shard, exchanging what dpba_solve_lm exchanges:
    the data terms of [H_pp | b_p | H_s | b_s] and the scalars (landmark energy, valid residuals, landmark norms) are
    summed over ranks with one all-reduce; priors, marginalised terms and the solve are replicated."""

    def __init__(self, O, frames, sigma, ab_reg):
        self.O, self.p = O, O.Problem(frames, sigma, ab_reg=ab_reg)

    @staticmethod
    def _sum(*arrays):
        flat = torch.from_numpy(np.concatenate([np.atleast_1d(np.asarray(a, dtype=np.float64)).ravel() for a in arrays]))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        out, k = [], 0
        for a in arrays:
            n = int(np.size(a))
            out.append(flat[k:k + n].numpy().reshape(np.shape(a)).copy())
            k += n
        return out

    def calculate_energy(self):
        O, p = self.O, self.p
        O.evaluate_jacobians(p.frames, p.sigma, fej=p.fej, evaluate_jacobians=False, new_point=True, huber=True)
        energy = 0.0
        for f in p.frames:  # replicated prior term
            ab = f.ab0 + f.state_eps[6:] + f.state_eps_step[6:]
            energy += 0.5 * float((ab * p.ab_reg) @ ab)
        le, nv = O.landmarks_energy(p.frames)
        (tot,) = self._sum(np.array([le, float(nv)]))
        return energy + float(tot[0]), int(round(tot[1]))

    def linearize(self):
        O, p = self.O, self.p
        O.evaluate_jacobians(p.frames, p.sigma, fej=p.fej, evaluate_jacobians=True, new_point=True, huber=True)
        Hp, bp = O.pose_pose(p.frames)
        Hs, bs = O.schur_complement(p.frames)
        p.H_pose, p.b_pose, p.H_schur, p.b_schur = self._sum(Hp, bp, Hs, bs)
        O.linear_system_prior(p.frames, p.H_pose, p.b_pose, p.ab_reg, p.fixed_reg)  # once, after the sum

    def calculate_step(self, lam):
        return self.p.calculate_step(lam)  # replicated solve + local back-substitution

    def accept_step(self):
        p = self.p
        lm_state = sum(float(f.idepth @ f.idepth) for f in p.frames)
        lm_step = sum(float(f.idepth_step @ f.idepth_step) for f in p.frames)
        state_sq, step_sq = p.accept_step()
        (tot,) = self._sum(np.array([lm_state, lm_step]))
        return state_sq - lm_state + float(tot[0]), step_sq - lm_step + float(tot[1])

    def reject_step(self):
        self.p.reject_step()

    def stop(self):
        return False


def lm_worker(rank, world, port, out):
    from oracle import pba_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    win = synth.make_window(n_frames=4, points_per_frame=101, seed=6, ab_scale=0.0)
    frames = O.frames_from_window(sharding.shard_window(win, rank, world))
    O.first_estimate_jacobians(frames)
    trace = []
    e, n, _ = O.lm_solve(ShardedProblem(O, frames, SIGMA, (1e12, 1e8)), O.LMOptions(7, 1e-5, 1e-8, 1e-8, True, 3, 1.0, 1.0), trace)
    np.save(out + f".state{rank}.npy", np.concatenate([O.state_eps_stacked(frames), [e, float(n), float(len(trace))]]))
    np.save(out + f".idepth{rank}.npy", frames[2].idepth)
    dist.destroy_process_group()


def test_sharded_lm_solve_equals_the_unsharded_solve(tmp_path):
    """Every rank takes the same accept / reject decisions from the same all-reduced sums, so the replicated frame
    state is identical on all ranks and equal to the single-process solve; landmark states gather back through the
    index map."""
    from oracle import pba_oracle as O
    world, out = 2, str(tmp_path / "lm")
    mp.spawn(lm_worker, args=(world, free_port(), out), nprocs=world, join=True)
    states = [np.load(out + f".state{r}.npy") for r in range(world)]
    assert np.array_equal(states[0], states[1])  # bit-identical replicas
    win = synth.make_window(n_frames=4, points_per_frame=101, seed=6, ab_scale=0.0)
    frames = O.frames_from_window(win)
    O.first_estimate_jacobians(frames)
    trace = []
    e, n, _ = O.lm_solve(O.Problem(frames, SIGMA), O.LMOptions(7, 1e-5, 1e-8, 1e-8, True, 3, 1.0, 1.0), trace)
    ref = np.concatenate([O.state_eps_stacked(frames), [e, float(n), float(len(trace))]])
    assert states[0][-1] == ref[-1] and states[0][-2] == ref[-2]  # iterations, valid residuals
    assert np.allclose(states[0][:-3], ref[:-3], rtol=0, atol=1e-10) and abs(states[0][-3] - ref[-3]) <= 1e-9 * abs(ref[-3])
    got = sharding.gather_landmark_array([np.load(out + f".idepth{r}.npy") for r in range(world)], 101, world)
    assert np.allclose(got, frames[2].idepth, rtol=0, atol=1e-10)
