"""Environment-sharded data parallelism of the update.

The reference has no multi-device code; correctness here is defined as "same result as the
single-device update on the union of the shards".  Rank r owns environments [lo, hi) of the
global rollout (all T steps): `compute_returns` needs no collective, advantage moments are
summed across ranks once per update, every rank draws the SAME global minibatch permutation
(same CPU seed) and processes the rows of each minibatch that it owns, loss terms are local
sums divided by the GLOBAL minibatch size, and the flat gradient buffer is all-reduced (sum)
before the clip + optimiser step, which every rank then performs identically.

This module is host logic only (pure index bookkeeping + the torch.distributed plumbing).
"""
import numpy as np
import torch


class ShardInfo(object):
    def __init__(self, world, rank, env_lo, env_hi, n_total):
        self.world, self.rank = int(world), int(rank)
        self.lo, self.hi, self.n_total = int(env_lo), int(env_hi), int(n_total)
        self.n_local = self.hi - self.lo


def shard_rows(global_rows, shard):
    """Rows of a global index list (flat index t * N_total + n) owned by this shard, renumbered
    to the local buffer (t * N_local + (n - lo)); order preserved."""
    g = np.asarray(global_rows, dtype=np.int64)
    t, n = g // shard.n_total, g % shard.n_total
    mine = (n >= shard.lo) & (n < shard.hi)
    return t[mine] * shard.n_local + (n[mine] - shard.lo)


def shard_epoch(perm, mbs_global, n_mb, shard, capacity=None):
    """Splits one epoch's permutation into n_mb minibatches and keeps this shard's rows.
    Returns (idx [n_mb, capacity] int64 padded with row 0, counts [n_mb])."""
    rows = [shard_rows(perm[k * mbs_global:(k + 1) * mbs_global], shard) for k in range(n_mb)]
    counts = np.array([len(r) for r in rows], dtype=np.int64)
    need = int(counts.max()) if len(rows) else 0
    if capacity is None or need > capacity:
        capacity = need
    idx = np.zeros((n_mb, max(capacity, 1)), dtype=np.int64)
    for k, r in enumerate(rows):
        idx[k, :len(r)] = r
    return idx, counts


def default_capacity(mbs_global, shard):
    """Rows per minibatch a shard must be able to hold: the expectation plus a 6-sigma
    hypergeometric margin, rounded up to a multiple of 4."""
    frac = shard.n_local / float(shard.n_total)
    mean = mbs_global * frac
    sigma = (mbs_global * frac * (1 - frac)) ** 0.5
    cap = int(mean + 6 * sigma + 4)
    return min(mbs_global, (cap + 3) // 4 * 4)


def shard_env_chunks(perm, E_global, n_mb, shard):
    """Recurrent minibatches of ONE global environment permutation (reference storage.py:152-156: chunk k =
    perm[k E : (k + 1) E]): the environments of each chunk that this shard owns, renumbered to the local buffer,
    in the order of the permutation.  -> list of n_mb int64 arrays (possibly empty)."""
    g = np.asarray(perm, dtype=np.int64)
    out = []
    for k in range(n_mb):
        c = g[k * E_global:(k + 1) * E_global]
        out.append(c[(c >= shard.lo) & (c < shard.hi)] - shard.lo)
    return out


def env_capacity(E_global, shard, gran):
    """Environments per recurrent minibatch a shard must be able to hold: the hypergeometric expectation plus a
    6-sigma margin, rounded up to the program granularity and capped by what the shard owns."""
    frac = shard.n_local / float(shard.n_total)
    mean = E_global * frac
    fpc = (shard.n_total - E_global) / max(shard.n_total - 1.0, 1.0)
    sigma = (E_global * frac * (1 - frac) * fpc) ** 0.5
    cap = int(mean + 6 * sigma + 1)
    cap = (cap + gran - 1) // gran * gran
    return max(gran, min(cap, (shard.n_local + gran - 1) // gran * gran))


def round_up_envs(count, gran):
    return max(gran, (int(count) + gran - 1) // gran * gran)


def all_reduce_sum(t):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
