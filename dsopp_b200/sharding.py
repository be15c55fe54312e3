"""Landmark sharding for the multi-GPU path.

Landmarks hosted by a keyframe are dealt round-robin over the ranks: global index l lives on rank l % world at
local index l // world.  All targets of a landmark stay on its rank (the per-landmark 1x1 Schur elimination and
the back-substitution are local); frames (images, masks, poses) are replicated.  The only exchange is the sum of
the packed reduced system [H_pp | b_p | H_s | b_s] per linearisation and of 8 scalars per energy evaluation.
"""
from __future__ import annotations

import copy

import numpy as np


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Global landmark indices owned by `rank` (ascending), i.e. l with l % world == rank."""
    return np.arange(rank, n, world, dtype=np.int64)


def to_local(l_global: np.ndarray, world: int):
    """global index -> (rank, local index)."""
    l_global = np.asarray(l_global, dtype=np.int64)
    return l_global % world, l_global // world


def to_global(rank, l_local, world: int):
    """(rank, local index) -> global index; exact inverse of to_local."""
    return np.asarray(l_local, dtype=np.int64) * world + np.asarray(rank, dtype=np.int64)


def shard_window(win, rank: int, world: int):
    """The sub-window a rank uploads: every frame, only this rank's landmarks and residual statuses."""
    out = copy.copy(win)
    out.frames = []
    idx = []
    for f in win.frames:
        s = shard_indices(len(f.idepth), rank, world)
        idx.append(s)
        g = copy.copy(f)
        g.uv, g.idepth, g.idepth_true, g.patch, g.flags = f.uv[s], f.idepth[s], f.idepth_true[s], f.patch[s], f.flags[s]
        out.frames.append(g)
    out.statuses = {(r, t): v[idx[r]] for (r, t), v in win.statuses.items()}
    return out


def gather_landmark_array(parts, n: int, world: int) -> np.ndarray:
    """Inverse of the deal: parts[rank] holds the values of that rank's landmarks in local order."""
    out = np.empty((n,) + np.asarray(parts[0]).shape[1:], dtype=np.asarray(parts[0]).dtype)
    for rank, p in enumerate(parts):
        out[shard_indices(n, rank, world)] = p
    return out
