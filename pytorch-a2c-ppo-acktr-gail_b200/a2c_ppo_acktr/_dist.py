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
import ctypes

import numpy as np
import torch

from . import _native

c_i32, c_i64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p

OP_COMM_FORK, OP_COMM_JOIN = 27, 28


class CommOpDesc(ctypes.Structure):
    """Mirror of a2cb_comm_op_t."""
    _fields_ = [("ctx", c_vp), ("buf", c_vp), ("count", c_i64)]


_native.register({
    "a2cb_nccl_load": (c_i32, [ctypes.c_char_p]),
    "a2cb_nccl_unique_id": (c_i32, [c_vp]),
    "a2cb_nccl_comm_init_rank": (c_i32, [ctypes.POINTER(c_vp), c_i32, c_i32, c_vp]),
    "a2cb_nccl_comm_destroy": (c_i32, [c_vp]),
    "a2cb_nccl_allreduce_sum_f32": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "a2cb_comm_ctx_create": (c_i32, [ctypes.POINTER(c_vp), c_vp]),
    "a2cb_comm_ctx_destroy": (c_i32, [c_vp]),
    "a2cb_comm_fork_allreduce": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "a2cb_comm_join": (c_i32, [c_vp, c_vp]),
    "a2cb_comm_ctx_set_p2p": (c_i32, [c_vp, c_vp]),
    "a2cb_p2p_create": (c_i32, [ctypes.POINTER(c_vp), c_i32, c_i32, c_i64]),
    "a2cb_p2p_ipc_handle": (c_i32, [c_vp, c_vp]),
    "a2cb_p2p_open_ipc": (c_i32, [c_vp, c_vp]),
    "a2cb_p2p_set_peers": (c_i32, [c_vp, ctypes.POINTER(c_vp)]),
    "a2cb_p2p_local_base": (c_i32, [c_vp, ctypes.POINTER(c_vp)]),
    "a2cb_p2p_capacity": (c_i64, [c_vp]),
    "a2cb_p2p_allreduce_sum_f32": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "a2cb_p2p_destroy": (c_i32, [c_vp]),
})


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


def balance_update_plan(perms, E_global, n_mb, world, n_local):
    """Reference-identical minibatches, balanced: who processes which environment of every global recurrent minibatch.

    `perms`: the `randperm(N_total)` of every epoch of one update (reference storage.py:152), identical on every rank;
    global minibatch (e, k) is perms[e][k E : (k + 1) E].  Its environments are spread over their owner ranks
    hypergeometrically; every rank keeps the first E / world of ITS environments (in chunk order), the rest are moved to
    the ranks that own fewer, donors and receivers matched in rank order -- the minimum number of moves, a pure function
    of the permutations, so every rank computes the same plan.  A moved environment lands in the next free GUEST slot of
    its receiver (column n_local + slot of the receiver's shadow rollout).

    Returns a list over ranks of dict(envs=[epoch][minibatch] -> int64 columns (E / world of them),
    send=[dst] -> local columns in transfer order, recv=[src] -> guest slots in the same order, n_guests,
    first_send=[dst] / first_recv=[src] -> how many leading entries of send / recv belong to the FIRST epoch)."""
    assert E_global % world == 0, "the global recurrent minibatch must split evenly over the ranks"
    # Vectorised over all minibatches of the update (the plan is made on the host while the device waits for it: ~0.15 ms
    # whatever the number of ranks; a loop over minibatches x ranks took 0.5 ms at 2 ranks, 2 ms at 8).
    target, E, n_ep = E_global // world, E_global, len(perms)
    M = n_ep * n_mb
    chunks = np.stack([np.asarray(p, dtype=np.int64)[:n_mb * E].reshape(n_mb, E) for p in perms]).reshape(M, E)
    owner = chunks // n_local
    order = np.argsort(owner.astype(np.uint8 if world <= 256 else np.int64), axis=1, kind="stable")   # (radix sort)
    order += np.arange(M)[:, None] * E                            # flat indices: chunk order inside every owner
    owner_s = owner.ravel()[order]
    local_s = chunks.ravel()[order] - owner_s * n_local
    counts = np.bincount((owner + np.arange(M)[:, None] * world).ravel(), minlength=M * world).reshape(M, world)
    starts = np.cumsum(counts, 1) - counts
    pos = np.arange(E)[None, :] - starts.ravel()[owner_s + np.arange(M)[:, None] * world]    # index inside the owner's group
    keep = pos < target
    need = np.maximum(target - counts, 0)                         # guests rank r takes in minibatch m
    base = np.cumsum(need, 0) - need                              # guest slots rank r has filled before minibatch m
    n_guests = need.sum(0)
    # surplus environments (minibatch-major, rank order inside a minibatch) pair up with the receivers' free places
    # (minibatch-major, rank order): donors and receivers matched in rank order
    s_m, s_p = np.nonzero(~keep)
    s_src, s_col = owner_s[s_m, s_p], local_s[s_m, s_p]
    r_m, r_rank = np.nonzero(need)
    reps = need[r_m, r_rank]
    s_dst = np.repeat(r_rank, reps)
    within = np.arange(int(reps.sum())) - np.repeat(np.cumsum(reps) - reps, reps)
    s_slot = np.repeat(base[r_m, r_rank], reps) + within
    assert len(s_dst) == len(s_src) and np.array_equal(np.repeat(r_m, reps), s_m)
    by_pair = np.argsort(s_src * world + s_dst, kind="stable")    # transfer order inside every (source, destination) pair
    pair_n = np.bincount(s_src * world + s_dst, minlength=world * world)
    first_n = np.bincount((s_src * world + s_dst)[s_m < n_mb], minlength=world * world)   # (stable order: they lead)
    cuts = np.concatenate([[0], np.cumsum(pair_n)]).tolist()      # pair (src, dst) = items cuts[src * world + dst] ...
    cols_by_pair, slots_by_pair = s_col[by_pair], s_slot[by_pair]
    envs = np.empty((world, M, target), np.int64)
    k_m, k_p = np.nonzero(keep)
    envs[owner_s[k_m, k_p], k_m, pos[k_m, k_p]] = local_s[k_m, k_p]          # kept environments, in chunk order
    envs[s_dst, s_m, np.minimum(counts[s_m, s_dst], target) + within] = n_local + s_slot
    out = []
    for r in range(world):
        out.append(dict(envs=[[envs[r, e * n_mb + k] for k in range(n_mb)] for e in range(n_ep)],
                        send=[cols_by_pair[cuts[r * world + d]:cuts[r * world + d + 1]] for d in range(world)],
                        recv=[slots_by_pair[cuts[s * world + r]:cuts[s * world + r + 1]] for s in range(world)],
                        n_guests=int(n_guests[r]),
                        first_send=[int(first_n[r * world + d]) for d in range(world)],
                        first_recv=[int(first_n[s * world + r]) for s in range(world)]))
    return out


class Exchange(object):
    """A variable all-to-all of fixed-size byte records (rows of `send` / `recv`, uint8 [n, record_bytes]): block d of
    `send` (send_counts[d] rows) goes to rank d, block s of `recv` arrives from rank s.  Yielded by the sharded update
    generators next to the tensors they want summed; `run` is the torch.distributed (NCCL) call, `emulate` performs
    the same movement between the objects of several emulated ranks in one process (tests)."""

    def __init__(self, send, send_counts, recv, recv_counts):
        self.send, self.recv = send, recv
        self.send_counts, self.recv_counts = [int(c) for c in send_counts], [int(c) for c in recv_counts]

    def run(self):
        import torch.distributed as dist
        dist.all_to_all_single(self.recv, self.send, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts)

    @staticmethod
    def emulate(items):
        so = [np.concatenate([[0], np.cumsum(x.send_counts)]) for x in items]
        ro = [np.concatenate([[0], np.cumsum(x.recv_counts)]) for x in items]
        for d, dst in enumerate(items):
            for s, src in enumerate(items):
                n = src.send_counts[d]
                assert n == dst.recv_counts[s]
                if n:
                    dst.recv[int(ro[d][s]):int(ro[d][s]) + n].copy_(src.send[int(so[s][d]):int(so[s][d]) + n])


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


class PeerExchange(object):
    """The peer-memory all-reduce of csrc/p2p.cu (a2cb_p2p_*): one exchange buffer per rank, mapped into the peers.
    `connect_ipc()` passes the CUDA IPC handles around through the default torch.distributed group (ranks = processes on
    one NVLink / NVSwitch node); `connect_local(group)` wires several instances living in ONE process (tests)."""

    def __init__(self, device, world, rank, capacity_floats):
        self.device = torch.device(device)
        self.world, self.rank = int(world), int(rank)
        h = c_vp()
        with torch.cuda.device(self.device):
            _native.check(_native.lib().a2cb_p2p_create(ctypes.byref(h), self.world, self.rank, int(capacity_floats)),
                          "a2cb_p2p_create")
        self.handle = h
        self.capacity = int(_native.lib().a2cb_p2p_capacity(h))

    def connect_ipc(self):
        import torch.distributed as dist
        lib = _native.lib()
        mine = torch.zeros(64, dtype=torch.uint8)
        _native.check(lib.a2cb_p2p_ipc_handle(self.handle, c_vp(mine.data_ptr())), "a2cb_p2p_ipc_handle")
        everyone = [None] * self.world
        dist.all_gather_object(everyone, bytes(mine.numpy().tobytes()))
        blob = torch.frombuffer(bytearray(b"".join(everyone)), dtype=torch.uint8)
        with torch.cuda.device(self.device):
            rc = lib.a2cb_p2p_open_ipc(self.handle, c_vp(blob.data_ptr()))
        # all ranks or none: a rank that cannot map its peers must not leave the others waiting for its arrivals
        ok = torch.tensor([1 if rc == 0 else 0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            raise _native.NativeError("a2cb_p2p_open_ipc: %s" % (_native.lib().a2cb_last_error().decode() if rc != 0 else "a peer could not map the buffers"))
        return self

    @staticmethod
    def connect_local(group):
        lib = _native.lib()
        bases = (c_vp * len(group))()
        for i, g in enumerate(group):
            b = c_vp()
            _native.check(lib.a2cb_p2p_local_base(g.handle, ctypes.byref(b)), "a2cb_p2p_local_base")
            bases[i] = b.value
        for g in group:
            _native.check(lib.a2cb_p2p_set_peers(g.handle, bases), "a2cb_p2p_set_peers")

    def all_reduce_sum(self, t, stream=None):
        """In-place fp32 sum of a flat, 16-byte aligned tensor on `stream` (default: the current stream)."""
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        sp = _native.stream_ptr(self.device) if stream is None else c_vp(stream.cuda_stream)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().a2cb_p2p_allreduce_sum_f32(self.handle, t.data_ptr(), t.numel(), sp),
                          "a2cb_p2p_allreduce_sum_f32")
        return t

    def close(self):
        if getattr(self, "handle", None):
            _native.lib().a2cb_p2p_destroy(self.handle)
            self.handle = None


class NativeComm(object):
    """An NCCL communicator owned by liba2cb (a2cb_nccl_* / a2cb_comm_* in include/a2cb.h) next to the default
    torch.distributed group: the 128-byte unique id is made on rank 0 and broadcast through the existing group, every
    rank then calls ncclCommInitRank.  `ctx` adds the side stream + events with which an op list overlaps the
    all-reduce of finished gradient segments with the remaining weight-gradient kernels (OP_COMM_FORK / OP_COMM_JOIN)."""

    def __init__(self, device):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise _native.NativeError("NativeComm needs an initialised torch.distributed process group for the rendezvous")
        lib = _native.lib()
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        try:                                          # the NCCL copy torch itself runs on
            import nvidia.nccl as n
            import os
            base = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                lib.a2cb_nccl_load(path.encode())
        except Exception:
            pass
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            _native.check(lib.a2cb_nccl_unique_id(c_vp(uid.data_ptr())), "a2cb_nccl_unique_id")
        uid_dev = uid.to(self.device)
        dist.broadcast(uid_dev, 0)
        uid = uid_dev.cpu()
        comm = c_vp()
        with torch.cuda.device(self.device):
            _native.check(lib.a2cb_nccl_comm_init_rank(ctypes.byref(comm), self.world, self.rank, c_vp(uid.data_ptr())),
                          "a2cb_nccl_comm_init_rank")
            ctx = c_vp()
            _native.check(lib.a2cb_comm_ctx_create(ctypes.byref(ctx), comm), "a2cb_comm_ctx_create")
        self.comm, self.ctx = comm, ctx

    def all_reduce_sum(self, t):
        """In-place fp32 sum on the current stream."""
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        with torch.cuda.device(self.device):
            _native.check(_native.lib().a2cb_nccl_allreduce_sum_f32(self.comm, t.data_ptr(), t.numel(),
                                                                    _native.stream_ptr(self.device)), "a2cb_nccl_allreduce_sum_f32")
        return t

    def enable_peer_exchange(self, capacity_floats):
        """Routes the forked segments through the peer-memory all-reduce (csrc/p2p.cu) instead of ncclAllReduce.
        Collective: every rank calls it at the same point (it is called while the op lists are built).  Falls back to
        NCCL -- on every rank -- when the buffers cannot be mapped; `A2CB_EXCHANGE=nccl` keeps NCCL."""
        import os
        import warnings
        if getattr(self, "p2p", None) is not None and self.p2p.capacity >= capacity_floats:
            return True
        if os.environ.get("A2CB_EXCHANGE", "p2p") == "nccl" or getattr(self, "_p2p_failed", False):
            return False
        try:
            p2p = PeerExchange(self.device, self.world, self.rank, capacity_floats).connect_ipc()
        except Exception as e:
            self._p2p_failed = True
            warnings.warn("peer-memory gradient exchange unavailable (%s): using ncclAllReduce" % e)
            return False
        old, self.p2p = getattr(self, "p2p", None), p2p
        _native.check(_native.lib().a2cb_comm_ctx_set_p2p(self.ctx, p2p.handle), "a2cb_comm_ctx_set_p2p")
        if old is not None:
            torch.cuda.synchronize(self.device)
            old.close()
        return True

    def fork_op(self, tensor, lo, hi):
        """Descriptor of an OP_COMM_FORK on tensor[lo:hi] (fp32, flat)."""
        d = CommOpDesc()
        d.ctx, d.buf, d.count = self.ctx, tensor.data_ptr() + 4 * lo, hi - lo
        return d

    def join_op(self):
        d = CommOpDesc()
        d.ctx = self.ctx
        return d

    def close(self):
        lib = _native.lib()
        if getattr(self, "ctx", None):
            lib.a2cb_comm_ctx_destroy(self.ctx)
            self.ctx = None
        if getattr(self, "p2p", None) is not None:
            self.p2p.close()
            self.p2p = None
        if getattr(self, "comm", None):
            lib.a2cb_nccl_comm_destroy(self.comm)
            self.comm = None


def native_comm(device):
    """NativeComm on `device` if the default process group runs NCCL on CUDA, else None (the gloo CPU tests and the
    single-GPU lock-step emulation keep exchanging through their own all-reduce of the yielded tensors)."""
    import torch.distributed as dist
    try:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl" \
                and torch.device(device).type == "cuda":
            return NativeComm(device)
    except Exception as e:                            # pragma: no cover - depends on the box
        import warnings
        warnings.warn("native NCCL communicator unavailable (%s): gradients are exchanged through torch.distributed" % e)
    return None


def gradient_buckets(layout):
    """[(lo, hi, label of the first backward op that must NOT precede the fork)] of the flat gradient buffer, in the
    order the backward pass finishes them: [GRU + heads (+ logstd)] once the GRU weight gradients are folded (before
    `gru_ih.dgrad`), [fc] once its weight / bias gradients are folded (before `fc.dgrad`), the convolutions at the end."""
    L = layout.layers
    total = layout.total
    if "fc" not in L:
        return [(0, total, None)]
    fc_lo = L["fc"][0]
    tail_lo = L["gru"][0] if "gru" in L else L["heads"][0]
    out = []
    if "gru" in L:
        out.append((tail_lo, total, "gru_ih.dgrad"))
        out.append((fc_lo, tail_lo, "fc.dgrad"))
    else:
        out.append((fc_lo, total, "fc.dgrad"))
    out.append((0, fc_lo, None))
    return out


def all_reduce_sum(t):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
