"""Interleaving model of the NVLink mailbox all-reduce protocol of dsopp_b200/csrc/peer_exchange.cu.

The kernel itself needs several GPUs; what can be checked on a CPU is the PROTOCOL: with two mailbox parities, one
exchange counter per rank (read by every CTA at start, bumped by the last CTA to finish) and per-(source, CTA) arrival
flags, does every rank always sum the W contributions of the CURRENT exchange -- for every interleaving of ranks and
CTAs, and for call sequences that mix grid sizes (17 CTAs for the system block, 1 for the scalar block)?  The model
executes the kernel's steps (read counter, push slice to each mailbox, raise flags, wait, sum, finish) as atomic
actions under a seeded random scheduler; kernels of one rank run in stream order, everything else is free to interleave.

A deliberately broken variant (parity taken from a per-CTA counter, the first design) must be caught by the same
model: that is the hazard the global counter removes.
"""
import random

import pytest

N_ELEM = 12  # elements of the modelled block; slices are contiguous ranges of it


class Rank:
    def __init__(self, world, n_ctas_max):
        self.seq = 0                      # completed exchanges
        self.done = 0                     # CTAs of the running kernel that finished
        self.cta_seq = [0] * n_ctas_max   # only used by the broken variant
        self.data = [[[None] * N_ELEM for _ in range(world)] for _ in range(2)]  # [parity][source][elem]
        self.flag = [[0] * n_ctas_max for _ in range(world)]                     # [source][cta]
        self.out = None


def cta_program(ranks, r, c, n_ctas, lo, hi, values, per_cta_parity):
    """Generator: one yield per atomic action of CTA c of rank r."""
    me = ranks[r]
    world = len(ranks)
    epoch = (me.cta_seq[c] if per_cta_parity else me.seq) + 1   # read at start
    par = epoch & 1
    yield
    for d in range(world):              # push the slice to every mailbox (posted stores: one action per destination)
        for i in range(lo, hi):
            ranks[d].data[par][r][i] = values[r][i]
        yield
    for d in range(world):              # fence + release store of the flag
        ranks[d].flag[r][c] = epoch
        yield
    for s in range(world):              # acquire loads until every source has arrived
        while me.flag[s][c] - epoch < 0:
            yield
    for i in range(lo, hi):             # rank-ordered sum from the own mailbox (one action per element: reads can
        me.out[i] = sum(me.data[par][s][i] for s in range(world))  # interleave with later pushes of faster ranks)
        yield
    if per_cta_parity:
        me.cta_seq[c] = epoch
    me.done += 1
    if me.done == n_ctas:               # the last CTA to finish publishes the epoch
        me.done = 0
        me.seq = epoch


def run(world, calls, seed, per_cta_parity=False):
    """calls: list of CTA counts, one per exchange.  Returns the number of wrong sums observed."""
    rng = random.Random(seed)
    n_max = max(calls)
    ranks = [Rank(world, n_max) for _ in range(world)]
    # contribution of rank r to element i in call k: a stale value (another call's) always changes the sum
    vals = [[[(k + 1) * 100000 + (r + 1) * 1009 + i * 131 for i in range(N_ELEM)] for r in range(world)]
            for k in range(len(calls))]
    call_of = [0] * world                 # exchange each rank is running
    live = [None] * world                 # generators of the running kernel per rank
    wrong = 0

    def start(r):
        k = call_of[r]
        n = calls[k]
        ranks[r].out = [None] * N_ELEM
        per = (N_ELEM + n - 1) // n
        live[r] = [cta_program(ranks, r, c, n, min(N_ELEM, c * per), min(N_ELEM, c * per + per), vals[k], per_cta_parity)
                   for c in range(n)]

    for r in range(world):
        start(r)
    steps = 0
    while any(g is not None for g in live):
        steps += 1
        assert steps < 2_000_000, "model does not terminate (deadlock)"
        r = rng.choice([q for q in range(world) if live[q] is not None])
        # a fast rank is favoured now and then so that it runs far ahead of the others
        if rng.random() < 0.3:
            r = min(q for q in range(world) if live[q] is not None)
        gens = live[r]
        g = rng.choice(gens)
        try:
            next(g)
        except StopIteration:
            gens.remove(g)
            if not gens:                  # kernel complete: check, then the next kernel of this rank may start
                k = call_of[r]
                expect = [sum(vals[k][s][i] for s in range(world)) for i in range(N_ELEM)]
                n = calls[k]
                per = (N_ELEM + n - 1) // n
                covered = min(N_ELEM, n * per)
                wrong += sum(1 for i in range(covered) if ranks[r].out[i] != expect[i])
                call_of[r] += 1
                if call_of[r] < len(calls):
                    start(r)
                else:
                    live[r] = None
    return wrong


CALLS = [4, 1, 4, 4, 1, 1, 4, 1, 4, 4, 4, 1, 4]  # system block / scalar block mixes, as dpba_solve_lm issues them


@pytest.mark.parametrize("world", [2, 3, 8])
def test_every_interleaving_sums_the_current_exchange(world):
    for seed in range(40):
        assert run(world, CALLS, seed) == 0


def test_model_catches_the_per_cta_parity_hazard():
    """With parity taken from a per-CTA counter the scalar-only calls advance CTA 0 alone, the halves used by different
    CTAs (and by different grid sizes for the same elements) drift apart, and a fast rank overwrites a slot a slow rank
    is still summing.  The model must see that -- otherwise it proves nothing about the global counter."""
    wrong = sum(run(2, CALLS, seed, per_cta_parity=True) for seed in range(200))
    assert wrong > 0
