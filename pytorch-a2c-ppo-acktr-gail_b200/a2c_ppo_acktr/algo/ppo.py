"""PPO update executed as fused op lists on the sm_100a engine.

Drop-in for `a2c_ppo_acktr.algo.PPO` (reference algo/ppo.py:8-96): same constructor,
`update(rollouts) -> (value_loss, action_loss, dist_entropy)` as Python floats, `.optimizer`
with writable `param_groups[i]['lr']`.  Per minibatch the reference dispatches ~200 small
torch kernels and three `.item()` syncs; here one C call launches

    gathered forward GEMMs -> loss epilogue -> backward GEMMs -> split-K folds -> clip + Adam

and the loss means come back with ONE device->host copy at the end of `update`.
Minibatch order: one `torch.randperm(T*N)` per epoch from the global CPU generator, chopped
into consecutive chunks with the remainder dropped -- the sampler semantics of
reference storage.py:122-125 -- so identical seeds give identical minibatches.
"""
import numpy as np
import torch

from .. import _dist, _native
from .optim import FlatAdam


def _adv_moments(rt, rollouts):
    """fp64 (sum, sum of squares, count) of returns - value_preds over this buffer's T*N rows."""
    T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
    scratch = rt.arena.ensure("adv_scratch", 2 * 128 + 2, torch.float64)
    moments = rt.arena.ensure("adv_moments", 4, torch.float64)
    with torch.cuda.device(rt.dev):
        _native.check(_native.lib().a2cb_adv_moments(
            _native.dptr(rollouts.returns, torch.float32, "returns"),
            _native.dptr(rollouts.value_preds, torch.float32, "value_preds"),
            T * N, scratch.data_ptr(), moments.data_ptr(), _native.stream_ptr(rt.dev)), "a2cb_adv_moments")
    return moments


def _adv_apply(rt, rollouts, moments):
    """(x - mean) / (unbiased std + 1e-5) from (possibly all-reduced) moments (ppo.py:35-37)."""
    T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
    adv = rt.arena.ensure("adv", T * N)
    with torch.cuda.device(rt.dev):
        _native.check(_native.lib().a2cb_adv_normalize(
            rollouts.returns.data_ptr(), rollouts.value_preds.data_ptr(), T * N, moments.data_ptr(),
            adv.data_ptr(), _native.stream_ptr(rt.dev)), "a2cb_adv_normalize")
    return adv


class _Phases(object):
    """A2CB_PHASE_TIMING=1: CUDA events and host clocks at named points of a sharded update; rank 0 prints, per phase, the
    device time since the previous mark and how far ahead of the device the host was when it enqueued the mark
    (negative lead = the device was waiting for the host)."""

    def __init__(self, dev, rank):
        import os
        self.on = os.environ.get("A2CB_PHASE_TIMING") == "1" and rank == 0
        self.dev, self.items = dev, []

    def mark(self, name):
        if self.on:
            import time
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.dev))
            self.items.append((name, ev, time.perf_counter()))

    def report(self):
        if not self.on or len(self.items) < 2:
            return
        import sys
        torch.cuda.synchronize(self.dev)
        n0, e0, h0 = self.items[0]
        prev = e0
        out = []
        for name, ev, h in self.items[1:]:
            out.append("%s %.3f (dev %.3f host %.3f)" % (name, prev.elapsed_time(ev), e0.elapsed_time(ev), (h - h0) * 1e3))
            prev = ev
        sys.stderr.write("phases [ms since previous mark (device clock, host clock since start)]: " + " | ".join(out) + "\n")


def _wait_staged(rollouts):
    f = getattr(rollouts, "wait_staged", None)
    if f is not None:
        f()


def _adv_normalize(rt, rollouts):
    return _adv_apply(rt, rollouts, _adv_moments(rt, rollouts))


class PPO():
    GRAPH_SLOTS = 4            # epochs whose index / scalar uploads may be in flight at once (_epoch_graph)
    ENV_GRANULARITY = 4        # sharded recurrent minibatches: local environment count rounded up to this many slots
    native_exchange = True     # sharded: exchange gradients inside the op list through liba2cb's NCCL communicator
    # sharded recurrent PPO.  "balanced" (default) and "global" both form the reference's minibatches -- the ONE
    # randperm(N_total) per epoch of storage.py:152.  "global": every rank processes the environments of a minibatch it
    # owns (hypergeometric counts: a step lasts as long as the fullest rank).  "balanced": the surplus environments of a
    # minibatch migrate to the ranks that own fewer (one all-to-all of rollout columns per update), so every rank runs
    # exactly E / world environments; it needs equal contiguous shards and E divisible by world, else "global" runs.
    # "stratified" = balanced per-rank permutations, NOT the reference's minibatch composition (round 1)
    recurrent_sharding = "balanced"

    def __init__(self,
                 actor_critic,
                 clip_param,
                 ppo_epoch,
                 num_mini_batch,
                 value_loss_coef,
                 entropy_coef,
                 lr=None,
                 eps=None,
                 max_grad_norm=None,
                 use_clipped_value_loss=True):
        self.actor_critic = actor_critic
        self.clip_param = clip_param
        self.ppo_epoch = ppo_epoch
        self.num_mini_batch = num_mini_batch
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.optimizer = FlatAdam(actor_critic, lr=lr, eps=eps)
        self._prog_key = None
        self._prog = None
        self.last_launches = 0
        self.step_hook = None      # optional callable(step_index) after each optimiser step (tests)
        self.shard = None          # _dist.ShardInfo when the rollout is sharded over ranks
        # One CUDA graph holds a whole epoch (num_mini_batch x ~20 kernels): the minibatch indices and the
        # per-step Adam scalars live in device buffers that are refreshed before each replay.
        self.use_cuda_graph = True
        self._graph = None

    def set_distributed(self, world, rank, env_range, n_total):
        """Declares that `rollouts` passed to update() hold environments [env_range) of a global
        rollout of n_total environments, one shard per rank of the default process group."""
        self.shard = _dist.ShardInfo(world, rank, env_range[0], env_range[1], n_total) if world > 1 else None
        self._prog_key = None
        self._rec_key = None
        self._comm, self._comm_tried = None, False
        if world > 1 and hasattr(self.actor_critic, "set_sampling_shard"):
            self.actor_critic.set_sampling_shard(env_range[0])      # exploration noise differs between ranks

    def _native_comm(self, rt):
        """liba2cb's own NCCL communicator (created on first use when the default process group runs NCCL): with it the
        gradient exchange is part of the op list -- finished segments are all-reduced on a side stream while the remaining
        weight-gradient kernels run (`_fuse_exchange`).  None: the sharded generators yield the gradient buffer and the
        caller all-reduces it between two op lists (torch.distributed, the lock-step tests, gloo)."""
        if self.shard is None or not self.native_exchange:
            return None
        if not getattr(self, "_comm_tried", False):
            self._comm_tried = True
            self._comm = _dist.native_comm(rt.dev)
        return getattr(self, "_comm", None)

    def _fuse_exchange(self, rt, prog, comm):
        """Turns a (forward + loss + backward) program into the whole sharded step: OP_COMM_FORK all-reduces of the
        gradient segments as the backward pass finishes them (GRU + heads, fc, convolutions: `_dist.gradient_buckets`),
        OP_COMM_JOIN, then grad-norm clip + Adam."""
        from .. import _engine
        G = rt.arena.bufs["G"]
        if hasattr(comm, "enable_peer_exchange"):
            comm.enable_peer_exchange(rt.L.total)                  # NVLink peer mappings instead of ncclAllReduce (csrc/p2p.cu)
        for lo, hi, before in _dist.gradient_buckets(rt.L):
            at = prog.index_of(before) if before is not None else None
            if at is None:
                at = len(prog.entries)
            prog.insert(at, _dist.OP_COMM_FORK, comm.fork_op(G, lo, hi), label="comm.fork")
        prog.add(_dist.OP_COMM_JOIN, comm.join_op(), label="comm.join")
        self.optimizer._ensure_state()
        prog.optim_desc = rt.add_optimizer_ops(prog, 0, self.optimizer.state1, self.optimizer.state2, self.max_grad_norm)
        prog.fused_exchange = True
        return prog.finalize()

    def _program(self, rt, rollouts, mbs, adv, accum, inv_B=None, seq=None, chunked=False):
        """(forward + loss + backward) program and the (clip + Adam) program.  Unsharded they are one
        op list; sharded they are separated by the gradient all-reduce.  `chunked`: the caller re-lays the
        frames out per minibatch during the first epoch (`_first_epoch_chunk`) when the program offers it."""
        key = (id(rt), rollouts.obs.data_ptr(), rollouts.actions.data_ptr(), rollouts.returns.data_ptr(),
               rollouts.value_preds.data_ptr(), rollouts.action_log_probs.data_ptr(), adv.data_ptr(),
               accum.data_ptr(), mbs, inv_B, seq, rollouts.masks.data_ptr(), self.clip_param, self.value_loss_coef, self.entropy_coef,
               self.use_clipped_value_loss, self.max_grad_norm)
        if key != self._prog_key:
            from .. import _engine
            hyper = dict(clip_param=self.clip_param, value_loss_coef=self.value_loss_coef,
                         entropy_coef=self.entropy_coef, use_clipped_value_loss=self.use_clipped_value_loss)
            if inv_B is not None:
                hyper["inv_B"] = inv_B
            prog = rt.train_program(rollouts, mbs, 0, hyper, adv=adv, loss_accum=accum, seq=seq)
            self.optimizer._ensure_state()
            comm = self._native_comm(rt)
            if comm is not None:
                opt = self._fuse_exchange(rt, prog, comm)          # one op list: exchange overlapped with the backward tail
            else:
                opt = prog if self.shard is None else _engine.Program(rt.dev)
                opt.optim_desc = rt.add_optimizer_ops(opt, 0, self.optimizer.state1, self.optimizer.state2,
                                                      self.max_grad_norm)
                prog.finalize()
                opt.finalize()
            self._prog, self._opt_prog, self._prog_key = prog, opt, key
        pro = getattr(self._prog, "prologue", None)
        if chunked and getattr(self._prog, "prologue_chunk", None) is not None:
            return self._prog, self._opt_prog
        _wait_staged(rollouts)
        if pro is not None and getattr(self, "_prologue_serial", None) != getattr(self, "_update_serial", 0):
            pro.run()                       # once per update: the rollout's frames do not change between minibatches
            self._prologue_serial = getattr(self, "_update_serial", 0)
        return self._prog, self._opt_prog

    @staticmethod
    def _first_epoch_chunks(prog, rollouts, perm, E, n_chunks):
        """First epoch of a recurrent update.  The frames of minibatch k's environments are re-laid out right before
        that minibatch (same bytes as one pass over the whole rollout); if the rollout was staged on the host
        (`RolloutStorage.stage_host`) their upload is started here, in minibatch order on the copy stream, and
        minibatch k only waits for ITS environments -- the rest of the upload overlaps the compute.
        Returns prepare(k, rows) to call before minibatch k, and finish() after the epoch."""
        progs = prog if isinstance(prog, (list, tuple)) else [prog] * n_chunks      # one program per chunk, or one for all
        chunks = perm if isinstance(perm, (list, tuple)) else [perm[k * E:(k + 1) * E] for k in range(n_chunks)]
        chunked = all(getattr(p, "prologue_chunk", None) is not None for p in progs)
        events = None
        if chunked and getattr(rollouts, "_staged_obs", None) is not None:
            events = rollouts._upload_staged(chunks)
        elif not chunked and isinstance(prog, (list, tuple)):
            # sized programs without a per-chunk re-layout: the whole rollout once (PPO._program does this for the
            # single-program paths)
            _wait_staged(rollouts)
            pro = getattr(progs[0], "prologue", None)
            if pro is not None:
                pro.run()
        cur = torch.cuda.current_stream(rollouts.rewards.device)

        def prepare(k, rows):
            if chunked:
                if events is not None:
                    cur.wait_event(events[k])
                progs[k].prologue_chunk.run(rows)

        def finish():
            if events is not None:
                for ev in events[n_chunks:]:
                    cur.wait_event(ev)
        return prepare, finish

    def update(self, rollouts):
        policy = self.actor_critic
        rt = policy.runtime()
        self._update_serial = getattr(self, "_update_serial", 0) + 1
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        if policy.is_recurrent:
            if self.shard is not None:
                return self._drive(self.sharded_recurrent_steps(rt, rollouts))
            return self._update_recurrent(rt, rollouts)
        batch_size = T * N
        assert batch_size >= self.num_mini_batch, (
            "PPO requires the number of processes ({}) "
            "* number of steps ({}) = {} "
            "to be greater than or equal to the number of PPO mini batches ({})."
            "".format(N, T, N * T, self.num_mini_batch))
        if self.shard is not None:
            return self._update_sharded(rt, rollouts)
        mbs = batch_size // self.num_mini_batch
        n_mb = batch_size // mbs

        adv = _adv_normalize(rt, rollouts)
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        prog, _ = self._program(rt, rollouts, mbs, adv, accum)
        launches0 = _native.launch_count()
        graphed = self.use_cuda_graph and self.step_hook is None
        for _ in range(self.ppo_epoch):
            perm = torch.randperm(batch_size)
            if graphed and self._epoch_graph(rt, prog, mbs, n_mb, perm):
                continue
            perm = perm.to(rt.dev)
            for k in range(n_mb):
                prog.optim_desc.dyn = None
                self.optimizer.configure(prog.optim_desc)
                prog.run(perm[k * mbs:(k + 1) * mbs])
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
        self.last_launches = _native.launch_count() - launches0
        num_updates = self.ppo_epoch * self.num_mini_batch
        sums = accum[:3].cpu()          # the update's only device->host synchronisation
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)

    def _epoch_graph(self, rt, prog, mbs, n_mb, perm):
        """Replays one epoch as a CUDA graph.  Returns False the first time a program is seen (that epoch
        runs eagerly, which also warms every kernel up); the graph is captured before the next one."""
        key = (id(prog), mbs, n_mb)
        g = self._graph
        if g is None or g["key"] != key:
            if g is not None and g.get("pending") == key:
                idx_buf = torch.zeros(n_mb * mbs, dtype=torch.int64, device=rt.dev)
                dyn_buf = torch.zeros(n_mb, 4, dtype=torch.float32, device=rt.dev)
                graph = torch.cuda.CUDAGraph()
                n0 = _native.launch_count()
                torch.cuda.synchronize(rt.dev)
                with torch.cuda.graph(graph):
                    for k in range(n_mb):
                        prog.optim_desc.dyn = dyn_buf.data_ptr() + 16 * k
                        prog.run(idx_buf[k * mbs:(k + 1) * mbs])
                prog.optim_desc.dyn = None
                # The uploads below are asynchronous and queue up behind the previous epoch's replay (milliseconds),
                # while the host only needs a randperm to reach the next epoch: every epoch in flight therefore needs
                # its OWN pinned staging slot.  A ring of slots, each guarded by an event recorded after its two
                # uploads; a slot is rewritten only after that event has completed.
                slots = [dict(dyn=torch.zeros(n_mb, 4).pin_memory(),
                              idx=torch.zeros(n_mb * mbs, dtype=torch.int64).pin_memory(), event=None)
                         for _ in range(self.GRAPH_SLOTS)]
                self._graph = g = dict(key=key, graph=graph, idx=idx_buf, dyn=dyn_buf,
                                       launches=_native.launch_count() - n0, slots=slots, next_slot=0)
                _native.lib().a2cb_add_launches(-g["launches"])          # capture recorded, did not run, them
            else:
                self._graph = dict(key=None, pending=key)
                return False
        opt = self.optimizer
        slot = g["slots"][g["next_slot"]]
        g["next_slot"] = (g["next_slot"] + 1) % len(g["slots"])
        if slot["event"] is not None:
            slot["event"].synchronize()          # the uploads that last read this slot have completed
        hd = slot["dyn"]
        for k in range(n_mb):
            opt.step_count += 1
            sc = opt.scalars()
            hd[k, 0], hd[k, 1], hd[k, 2] = sc["lr_eff"], sc["bc2_sqrt"], float(sc["first_step"])
        slot["idx"].copy_(perm[:n_mb * mbs])
        g["idx"].copy_(slot["idx"], non_blocking=True)
        g["dyn"].copy_(hd, non_blocking=True)
        if slot["event"] is None:
            slot["event"] = torch.cuda.Event()
        slot["event"].record(torch.cuda.current_stream(rt.dev))
        g["graph"].replay()
        _native.lib().a2cb_add_launches(g["launches"])
        return True

    def _update_recurrent(self, rt, rollouts):
        """Recurrent minibatches are whole environments (reference storage.py:145-202): one
        randperm(N) per epoch, chunks of E = N // num_mini_batch environments, rows time-major."""
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        assert N >= self.num_mini_batch, (
            "PPO requires the number of processes ({}) "
            "to be greater than or equal to the number of "
            "PPO mini batches ({}).".format(N, self.num_mini_batch))
        E = N // self.num_mini_batch
        ph = _Phases(rt.dev, 0)
        ph.mark("start")
        adv = _adv_normalize(rt, rollouts)
        ph.mark("adv")
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        prog, _ = self._program(rt, rollouts, T * E, adv, accum, seq=(T, E), chunked=True)
        hxs_mb = rt.arena.bufs["hxs_mb"][:E * rt.spec.hidden].view(E, rt.spec.hidden)
        h0 = rollouts.recurrent_hidden_states[0]
        t_off = (torch.arange(T, device=rt.dev, dtype=torch.int64) * N).view(T, 1)
        launches0 = _native.launch_count()
        for epoch in range(self.ppo_epoch):
            perm = torch.randperm(N)
            if epoch == 0:
                prepare, finish = self._first_epoch_chunks(prog, rollouts, perm, E, N // E)
            for k, start in enumerate(range(0, N, E)):
                if start + E > N:
                    raise IndexError("index {} is out of bounds for dimension 0 with size {}".format(N, N))
                env = perm[start:start + E].to(rt.dev)
                rows = (t_off + env.view(1, E)).reshape(-1).contiguous()
                if epoch == 0:
                    prepare(k, rows)
                _native.gather_rows([(h0, hxs_mb)], env, E)
                self.optimizer.configure(prog.optim_desc)
                prog.run(rows)
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
                ph.mark("mb%d.%d" % (epoch, k))
            if epoch == 0:
                finish()
        self.last_launches = _native.launch_count() - launches0
        num_updates = self.ppo_epoch * self.num_mini_batch
        sums = accum[:3].cpu()
        ph.mark("losses")
        ph.report()
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)

    def _update_sharded(self, rt, rollouts):
        return self._drive(self.sharded_steps(rt, rollouts))

    @staticmethod
    def _drive(gen):
        try:
            while True:
                item = next(gen)
                if isinstance(item, _dist.Exchange):
                    item.run()
                else:
                    _dist.all_reduce_sum(item)
        except StopIteration as done:
            return done.value

    def sharded_recurrent_steps(self, rt, rollouts):
        """Recurrent PPO on an environment shard; yields the tensors to sum-all-reduce, like `sharded_steps`.
        `recurrent_sharding` selects how the global minibatches are formed (see the two methods)."""
        if self.recurrent_sharding == "stratified":
            return self._sharded_recurrent_stratified(rt, rollouts)
        sh = self.shard
        nmb = self.num_mini_batch
        if (self.recurrent_sharding == "balanced" and sh.world > 1 and sh.n_local * sh.world == sh.n_total
                and sh.lo == sh.rank * sh.n_local and sh.n_total % (nmb * sh.world) == 0):
            return self._sharded_recurrent_balanced(rt, rollouts)
        return self._sharded_recurrent_global(rt, rollouts)

    def _sized_recurrent_program(self, rt, rollouts, E_p, adv, accum, inv_B):
        """Forward + loss + backward program for E_p environment slots (rows time-major, slot e valid iff
        e < n_valid), and the shared clip + Adam program.  All sizes share one set of activation buffers (tag "tR_"):
        the largest size is built first and the cache is dropped if a larger one is ever needed."""
        from .. import _engine
        T = rollouts.rewards.shape[0]
        key = (id(rt), rollouts.obs.data_ptr(), rollouts.actions.data_ptr(), rollouts.returns.data_ptr(),
               rollouts.value_preds.data_ptr(), rollouts.action_log_probs.data_ptr(), adv.data_ptr(), accum.data_ptr(),
               inv_B, rollouts.masks.data_ptr(), self.clip_param, self.value_loss_coef, self.entropy_coef,
               self.use_clipped_value_loss, self.max_grad_norm)
        if key != getattr(self, "_rec_key", None):
            self._rec_key, self._rec_progs, self._rec_opt = key, {}, None
        prog = self._rec_progs.get(E_p)
        if prog is None:
            if self._rec_progs and E_p > max(self._rec_progs):
                self._rec_progs = {}
            hyper = dict(clip_param=self.clip_param, value_loss_coef=self.value_loss_coef, entropy_coef=self.entropy_coef,
                         use_clipped_value_loss=self.use_clipped_value_loss, inv_B=inv_B)
            prog = rt.train_program(rollouts, T * E_p, 0, hyper, adv=adv, loss_accum=accum, seq=(T, E_p), tag="tR_")
            prog.loss_desc.valid_mod = E_p
            comm = self._native_comm(rt)
            if comm is not None:
                self._fuse_exchange(rt, prog, comm)
            else:
                prog.finalize()
            self._rec_progs[E_p] = prog
        if getattr(prog, "fused_exchange", False):
            return prog, prog
        if self._rec_opt is None:
            self.optimizer._ensure_state()
            opt = _engine.Program(rt.dev)
            opt.optim_desc = rt.add_optimizer_ops(opt, 0, self.optimizer.state1, self.optimizer.state2, self.max_grad_norm)
            self._rec_opt = opt.finalize()
        return prog, self._rec_opt

    def _sharded_recurrent_global(self, rt, rollouts):
        """Reference-identical minibatches (the default): every rank draws the ONE `randperm(N_total)` of reference
        storage.py:152 (same CPU seed), global minibatch k is perm[k E : (k + 1) E], and each rank processes the
        environments of that chunk it owns -- no frame or hidden state ever leaves its GPU.  The local count is
        hypergeometric, so the program runs on the count rounded up to `ENV_GRANULARITY` environment slots; padding
        slots repeat a real environment and contribute neither loss nor gradient (a2cb_loss_t.n_valid / valid_mod);
        loss terms are local sums over the GLOBAL minibatch size."""
        sh = self.shard
        self._update_serial = getattr(self, "_update_serial", 0) + 1
        T, n_loc = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        assert n_loc == sh.n_local, "rollouts hold %d environments, the shard declares %d" % (n_loc, sh.n_local)
        nmb = self.num_mini_batch
        assert sh.n_total >= nmb and sh.n_total % nmb == 0, (
            "sharded recurrent PPO needs a number of processes ({}) divisible by the number of PPO mini batches ({})"
            .format(sh.n_total, nmb))
        E_g = sh.n_total // nmb
        gran = self.ENV_GRANULARITY if n_loc >= 2 * self.ENV_GRANULARITY else 1
        cap = _dist.env_capacity(E_g, sh, gran)
        moments = _adv_moments(rt, rollouts)
        yield moments
        adv = _adv_apply(rt, rollouts, moments)
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        inv_B = 1.0 / (T * E_g)
        self._sized_recurrent_program(rt, rollouts, cap, adv, accum, inv_B)        # largest first: owns the buffers
        H = rt.spec.hidden
        h0 = rollouts.recurrent_hidden_states[0]
        t_off = (torch.arange(T, device=rt.dev, dtype=torch.int64) * n_loc).view(T, 1)
        grads = rt.arena.bufs["G"]
        launches0 = _native.launch_count()
        for epoch in range(self.ppo_epoch):
            perm = torch.randperm(sh.n_total).numpy()         # identical on every rank (same CPU seed)
            chunks = _dist.shard_env_chunks(perm, E_g, nmb, sh)
            plan = []
            for mine in chunks:
                E_p = _dist.round_up_envs(len(mine), gran)
                prog, opt = self._sized_recurrent_program(rt, rollouts, E_p, adv, accum, inv_B)
                env = torch.full((E_p,), int(mine[0]) if len(mine) else 0, dtype=torch.int64)
                env[:len(mine)] = torch.from_numpy(mine)
                plan.append((prog, env, len(mine), E_p))
            if epoch == 0:
                prepare, finish = self._first_epoch_chunks([p[0] for p in plan], rollouts,
                                                           [torch.from_numpy(c) for c in chunks], None, nmb)
            for k, (prog, env, count, E_p) in enumerate(plan):
                env = env.to(rt.dev)
                rows = (t_off + env.view(1, E_p)).reshape(-1).contiguous()
                if epoch == 0:
                    prepare(k, rows)
                hxs_mb = rt.arena.bufs["hxs_mb"][:E_p * H].view(E_p, H)
                _native.gather_rows([(h0, hxs_mb)], env, E_p)
                prog.loss_desc.n_valid = count
                if getattr(prog, "fused_exchange", False):
                    self.optimizer.configure(prog.optim_desc)
                    prog.run(rows)                                  # exchange + clip + Adam are part of the op list
                else:
                    prog.run(rows)
                    yield grads
                    self.optimizer.configure(opt.optim_desc)
                    opt.run()
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
            if epoch == 0:
                finish()
        self.last_launches = _native.launch_count() - launches0
        yield accum
        num_updates = self.ppo_epoch * self.num_mini_batch
        sums = accum[:3].cpu()
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)

    # ---- reference-identical minibatches, balanced by migrating environments -------------------------------------
    _MIGRATED = ("masks", "value_preds", "returns", "action_log_probs", "actions")

    def _shadow_rollout(self, rt, rollouts, n_guests):
        """The rollout this rank trains on in balanced mode: its own environments (columns 0 .. n_local - 1, copied
        from `rollouts` at the start of every update) followed by guest columns for migrated environments."""
        from ..storage import RolloutStorage
        T, n_loc = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        cap = getattr(self, "_shadow_cap", 0)
        if n_guests > cap or getattr(self, "_shadow_src", None) != (rollouts.rewards.data_ptr(), T, n_loc):
            cap = max(cap, 96, (int(n_guests * 1.25) + 31) // 32 * 32)
            ring = bool(getattr(rollouts, "frame_ring", False))
            obs_shape = tuple(rollouts.obs.shape[2:])

            # (the storage duck-types its action space by class name, like the reference: storage.py:19,24)
            space = type("Discrete" if rollouts.actions.dtype == torch.int64 else "Box", (object,),
                         {"shape": (rollouts.actions.shape[-1],), "n": 0})()
            # (+ 1: a dummy guest column that the fixed-shape scatter of unused message rows writes to)
            sh = RolloutStorage(T, n_loc + cap + 1, obs_shape, space, rollouts.recurrent_hidden_states.shape[-1],
                                obs_dtype=torch.uint8 if ring else rollouts.obs.dtype, frame_ring=ring)
            sh.to(rt.dev)
            self._shadow, self._shadow_cap = sh, cap
            self._shadow_src = (rollouts.rewards.data_ptr(), T, n_loc)
        return self._shadow

    @staticmethod
    def _relayout_columns(prog, cols, T, E, n_columns, dev):
        """Frame re-layout (`prologue_chunk`, built for E columns = T * E images) of the rollout columns `cols` (host
        int64), at most E per launch; the image count is a field of the op's descriptor."""
        if len(cols) == 0:
            return
        s2d = prog.prologue_chunk.entries[0][2]
        full = s2d.B
        assert len(prog.prologue_chunk.entries) == 1 and full == T * E
        c = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int64)).to(dev)
        rows = (torch.arange(T, device=dev, dtype=torch.int64).view(T, 1) * n_columns + c.view(1, -1))
        try:
            for i in range(0, len(cols), E):
                part = rows[:, i:i + E].reshape(-1).contiguous()
                s2d.B = part.numel()
                prog.prologue_chunk.run(part)
        finally:
            s2d.B = full

    def _migration_fields(self, st):
        """(tensor, leading rows) of everything an update reads per environment (dimension 1 = environment)."""
        f = [(st.frames, None), (st.stack_valid, None)] if getattr(st, "frame_ring", False) else [(st.obs, None)]
        f += [(getattr(st, k), None) for k in self._MIGRATED]
        f.append((st.recurrent_hidden_states, 1))              # only the initial state is read (storage.py:160-161)
        return f

    def _sharded_recurrent_balanced(self, rt, rollouts):
        """Reference-identical minibatches (the ONE `randperm(N_total)` per epoch of reference storage.py:152) with every
        rank processing exactly E / world environments of each: `_dist.balance_update_plan` moves the surplus
        environments of a minibatch to the ranks that own fewer.  All permutations of the update are drawn first (the
        same RNG sequence as drawing one per epoch), the rollout columns of every environment that has to move travel in
        ONE all-to-all per update (~0.9 MB per environment in frame-ring format, a few dozen per rank), and the update
        then runs on a shadow rollout [own environments | guests] with a single program size.

        Frames staged on the host (`stage_host`) arrive in the order they are needed -- the environments the first epoch
        sends away, the ones its first minibatch keeps (both known from the first permutation, so the copy engine starts
        before the plan is made), then minibatch by minibatch -- and the migration runs in TWO all-to-alls: what the first
        epoch needs before it (while the first minibatch's own columns are still arriving), the later epochs' guests after
        it (by then every column is on the device)."""
        sh = self.shard
        self._update_serial = getattr(self, "_update_serial", 0) + 1
        T, n_loc = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        assert n_loc == sh.n_local, "rollouts hold %d environments, the shard declares %d" % (n_loc, sh.n_local)
        nmb, n_ep = self.num_mini_batch, self.ppo_epoch
        E_g = sh.n_total // nmb
        E = E_g // sh.world
        ph = _Phases(rt.dev, sh.rank)
        ph.mark("start")
        dev = rt.dev
        cur = torch.cuda.current_stream(dev)
        moments = _adv_moments(rt, rollouts)                    # (device work first: the host-side planning overlaps it)
        upload = rollouts._staged_upload() if hasattr(rollouts, "_staged_upload") else None
        staged = upload is not None
        src_f = self._migration_fields(rollouts)

        def own_columns(shadow):
            """The rank's own environments -> columns [0, n_loc) of the shadow rollout (frames staged on the host follow
            their upload, chunk by chunk)."""
            pairs = [((b if rows is None else b[:rows])[:, :n_loc], a if rows is None else a[:rows])
                     for (a, rows), (b, _) in zip(src_f[1 if staged else 0:], self._migration_fields(shadow)[1 if staged else 0:])]
            _native.copy_columns(pairs, None, None, n_loc)          # one launch (a strided uint8 copy_ per field is slow)
        early = getattr(self, "_shadow", None) if getattr(self, "_shadow_src", None) == (rollouts.rewards.data_ptr(), T, n_loc) else None
        pre = None
        if early is not None:
            own_columns(early)
            bp = getattr(self, "_bal_prog", None)
            if (not staged and bp is not None and bp[0] is early and getattr(bp[1], "prologue_chunk", None) is not None
                    and n_ep > 1):
                # resident frames: every own environment is used by this update (locally or as somebody's guest), so
                # all own columns are re-laid out NOW -- device work that does not depend on the plan the host is about
                # to make -- and the first epoch's minibatches find their frames ready
                self._relayout_columns(bp[1], np.arange(n_loc, dtype=np.int64), T, E, early.rewards.shape[1], dev)
                pre = bp[1]
        perms = [torch.randperm(sh.n_total).numpy() for _ in range(n_ep)]
        events = {}
        if staged:
            # what the first epoch sends away and what its first minibatch keeps follow from the first permutation
            # alone (a rank keeps the first E of its environments of a chunk, in chunk order): the copy engine gets
            # them -- in that order: the all-to-all can then run while the kept columns are still arriving -- before
            # the host makes the full plan
            p0 = np.asarray(perms[0], dtype=np.int64)
            own_k = [c[c // n_loc == sh.rank] - sh.rank * n_loc for c in (p0[k * E_g:(k + 1) * E_g] for k in range(nmb))]
            sentA = np.unique(np.concatenate([c[E:] for c in own_k]))
            kept0 = np.sort(own_k[0][:E])
            events["sentA"] = upload.columns(sentA)
            events["kept0"] = upload.columns(kept0)
        plans = _dist.balance_update_plan(perms, E_g, nmb, sh.world, n_loc)
        mine = plans[sh.rank]
        ph.mark("plan")
        yield moments
        ph.mark("moments")
        shadow = self._shadow_rollout(rt, rollouts, max(p["n_guests"] for p in plans))
        if shadow is not early:
            own_columns(shadow)
        Nv = shadow.rewards.shape[1]
        dst_f = self._migration_fields(shadow)
        big_src, big_dst = src_f[0][0], dst_f[0][0]              # frames (ring) / obs: everything else is narrow
        # two migration phases for staged frames (first epoch / later epochs), one otherwise
        two = staged and n_ep > 1
        cutS = [int(c) for c in mine["first_send"]] if two else [len(x) for x in mine["send"]]
        cutR = [int(c) for c in mine["first_recv"]] if two else [len(x) for x in mine["recv"]]
        parts = [([x[:c] for x, c in zip(mine["send"], cutS)], [x[:c] for x, c in zip(mine["recv"], cutR)])]
        if two:
            parts.append(([x[c:] for x, c in zip(mine["send"], cutS)], [x[c:] for x, c in zip(mine["recv"], cutR)]))
        moves_all = sum(len(x) for p in plans for x in p["send"])
        moves_first = sum(sum(p["first_send"]) for p in plans) if two else moves_all
        moves = [moves_first, moves_all - moves_first]          # over ALL ranks: a phase nobody takes part in is skipped
        first_cols = [x[x < n_loc] for x in mine["envs"][0]]
        if staged:
            assert np.array_equal(sentA, np.unique(np.concatenate(
                [x[:c] for x, c in zip(mine["send"], mine["first_send"])] + [np.zeros(0, np.int64)])))
            assert np.array_equal(kept0, np.sort(first_cols[0]))
            for k in range(1, nmb):
                events[k] = upload.columns(first_cols[k])
            events["rest"] = upload.finish()                    # none are left by construction; closes the upload (+ tail)

        def frames_to_shadow(cols, event):
            cur.wait_event(event)
            if len(cols):
                c = torch.from_numpy(np.ascontiguousarray(cols)).to(dev)
                _native.copy_columns([(big_dst, big_src)], c, c, len(cols))
        if staged:
            frames_to_shadow(sentA, events["sentA"])
        # ---- migration: the travelling environments' columns packed into records (one record = every field of one
        #      environment, environment-major), an all-to-all, records unpacked into the guest columns.  The record
        #      buffers have the CAPACITY of the shadow rollout (no allocator growth, one registration), the kernels and
        #      the messages carry this update's counts. ----
        cap = Nv - n_loc - 1
        n_send, n_recv = sum(len(x) for x in mine["send"]), sum(len(x) for x in mine["recv"])
        assert n_send <= cap and n_recv <= cap
        cols_np = np.zeros(2 * cap, np.int64)
        so, ro, bounds = 0, 0, []
        for snd, rcv in parts:                                  # phase-major, destination / source rank order inside
            a, b = np.concatenate(snd + [np.zeros(0, np.int64)]), np.concatenate(rcv + [np.zeros(0, np.int64)])
            cols_np[so:so + len(a)] = a
            cols_np[cap + ro:cap + ro + len(b)] = b + n_loc
            bounds.append((so, so + len(a), ro, ro + len(b)))
            so, ro = so + len(a), ro + len(b)
        cols_dev = torch.from_numpy(cols_np).to(dev)
        views = [(x if rows is None else x[:rows]) for x, rows in src_f]
        dviews = [(x if rows is None else x[:rows]) for x, rows in dst_f]
        widths = [(v[:, 0].numel() * v.element_size() + 15) // 16 * 16 for v in views]      # fields start 16-byte aligned
        rec_bytes = sum(widths)
        bufs = getattr(self, "_xchg_bufs", None)
        if bufs is None or bufs[0].shape != (cap, rec_bytes) or bufs[0].device != dev:
            bufs = self._xchg_bufs = (torch.zeros(cap, rec_bytes, dtype=torch.uint8, device=dev),
                                      torch.zeros(cap, rec_bytes, dtype=torch.uint8, device=dev))
        send, recv = bufs

        def record_views(buf, like):
            """The fields of `like` (tensors [rows, columns, ...]) as [rows, record, bytes] views of the record buffer."""
            out, off = [], 0
            for v, w in zip(like, widths):
                rows, wb = v.shape[0], v[0, 0].numel() * v.element_size()
                out.append(buf.narrow(1, off, rows * wb).unflatten(1, (rows, wb)).transpose(0, 1))
                off += w
            return out

        def migrate(i):
            """Phase i of the migration (a generator: the all-to-all is performed by whoever drives the update)."""
            s0, s1, r0, r1 = bounds[i]
            snd, rcv = parts[i]
            if moves[i] == 0:
                return
            _native.copy_columns(list(zip(record_views(send[s0:s1], views), views)), None, cols_dev[s0:], s1 - s0)
            ph.mark("pack%d" % i)
            yield _dist.Exchange(send[s0:s1], [len(x) for x in snd], recv[r0:r1], [len(x) for x in rcv])
            ph.mark("exchange%d" % i)
            _native.copy_columns(list(zip(dviews, record_views(recv[r0:r1], dviews))), cols_dev[cap + r0:], None, r1 - r0)
        yield from migrate(0)
        adv = _adv_apply(rt, shadow, moments)
        ph.mark("unpack+adv")
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        inv_B = 1.0 / (T * E_g)
        prog, opt = self._sized_recurrent_program(rt, shadow, E, adv, accum, inv_B)
        H = rt.spec.hidden
        h0 = shadow.recurrent_hidden_states[0]
        t_off = (torch.arange(T, device=dev, dtype=torch.int64) * Nv).view(T, 1)
        grads = rt.arena.bufs["G"]
        launches0 = _native.launch_count()
        chunked = getattr(prog, "prologue_chunk", None) is not None
        if not chunked and getattr(prog, "prologue", None) is not None:
            assert not two, "staged frames need the per-chunk re-layout"
            prog.prologue.run()
        self._bal_prog = (shadow, prog)
        pre_laid = chunked and pre is prog                      # (a rebuilt program re-lays everything out below)
        envs = torch.from_numpy(np.concatenate([x for ep in mine["envs"] for x in ep])).to(dev).view(n_ep, nmb, E)
        rows_all = (t_off.view(1, 1, T, 1) + envs.view(n_ep, nmb, 1, E)).reshape(n_ep, nmb, T * E)
        later_guests = np.zeros(0, np.int64)
        if pre_laid:
            self._relayout_columns(prog, n_loc + np.arange(n_recv, dtype=np.int64), T, E, Nv, dev)      # the guests
        elif chunked and n_ep > 1:
            # columns first used after the first epoch (own environments that the first epoch sent away, guests of later
            # epochs): their frames are re-laid out now -- the guests of a second migration phase once they have arrived
            first = np.unique(np.concatenate(mine["envs"][0]))
            later = np.setdiff1d(np.unique(np.concatenate([x for ep in mine["envs"][1:] for x in ep])), first)
            if two:
                later, later_guests = later[later < n_loc], later[later >= n_loc]
            self._relayout_columns(prog, later, T, E, Nv, dev)
        ph.mark("relayout-later")
        for epoch in range(n_ep):
            if epoch == 1 and two:
                yield from migrate(1)
                assert _adv_apply(rt, shadow, moments).data_ptr() == adv.data_ptr()     # the new guests' advantages
                self._relayout_columns(prog, later_guests, T, E, Nv, dev)
                ph.mark("migrate1")
            for k in range(nmb):
                env, rows = envs[epoch, k], rows_all[epoch, k]
                if epoch == 0 and staged:
                    frames_to_shadow(first_cols[k], events["kept0" if k == 0 else k])
                    if k == nmb - 1:
                        cur.wait_event(events["rest"])
                if epoch == 0 and chunked and not pre_laid:
                    prog.prologue_chunk.run(rows)
                hxs_mb = rt.arena.bufs["hxs_mb"][:E * H].view(E, H)
                _native.gather_rows([(h0, hxs_mb)], env, E)
                prog.loss_desc.n_valid = E
                if getattr(prog, "fused_exchange", False):
                    self.optimizer.configure(prog.optim_desc)
                    prog.run(rows)
                else:
                    prog.run(rows)
                    yield grads
                    self.optimizer.configure(opt.optim_desc)
                    opt.run()
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
                ph.mark("mb%d.%d" % (epoch, k))
        self.last_launches = _native.launch_count() - launches0
        yield accum
        num_updates = n_ep * self.num_mini_batch
        sums = accum[:3].cpu()
        ph.mark("losses")
        ph.report()
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)

    def _sharded_recurrent_stratified(self, rt, rollouts):
        """Opt-in (`recurrent_sharding = "stratified"`): global minibatch k is the union over ranks of the k-th chunk
        of each rank's LOCAL environment permutation -- perfectly balanced, nothing padded, but NOT the reference's
        minibatch composition (the partition is stratified by rank instead of one global randperm).  A single-device
        run fed the same environment sets gives the same result."""
        sh = self.shard
        self._update_serial = getattr(self, "_update_serial", 0) + 1
        T, n_loc = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        assert n_loc == sh.n_local and n_loc >= self.num_mini_batch
        E = n_loc // self.num_mini_batch
        E_global = E * sh.world
        ph = _Phases(rt.dev, sh.rank)
        ph.mark("start")
        moments = _adv_moments(rt, rollouts)
        yield moments
        adv = _adv_apply(rt, rollouts, moments)
        ph.mark("moments+adv")
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        prog, opt = self._program(rt, rollouts, T * E, adv, accum, inv_B=1.0 / (T * E_global), seq=(T, E), chunked=True)
        hxs_mb = rt.arena.bufs["hxs_mb"][:E * rt.spec.hidden].view(E, rt.spec.hidden)
        h0 = rollouts.recurrent_hidden_states[0]
        t_off = (torch.arange(T, device=rt.dev, dtype=torch.int64) * n_loc).view(T, 1)
        grads = rt.arena.bufs["G"]
        launches0 = _native.launch_count()
        for epoch in range(self.ppo_epoch):
            perm = torch.randperm(n_loc)                      # same seed on every rank -> same local order
            if epoch == 0:
                prepare, finish = self._first_epoch_chunks(prog, rollouts, perm, E, self.num_mini_batch)
            for k, start in enumerate(range(0, E * self.num_mini_batch, E)):
                env = perm[start:start + E].to(rt.dev)
                rows = (t_off + env.view(1, E)).reshape(-1).contiguous()
                if epoch == 0:
                    prepare(k, rows)
                _native.gather_rows([(h0, hxs_mb)], env, E)
                if getattr(prog, "fused_exchange", False):
                    self.optimizer.configure(prog.optim_desc)
                    prog.run(rows)                                  # exchange + clip + Adam are part of the op list
                else:
                    prog.run(rows)
                    yield grads
                    self.optimizer.configure(opt.optim_desc)
                    opt.run()
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
                ph.mark("mb%d.%d" % (epoch, k))
            if epoch == 0:
                finish()
        self.last_launches = _native.launch_count() - launches0
        yield accum
        ph.mark("accum")
        ph.report()
        num_updates = self.ppo_epoch * self.num_mini_batch
        sums = accum[:3].cpu()
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)

    def sharded_steps(self, rt, rollouts):
        """Same update on an environment shard (see _dist.py): global permutation, own rows only,
        gradient all-reduce before the (redundant, identical) clip + Adam step.  Written as a
        generator that yields every tensor that must be sum-all-reduced in place, so that the
        collective can be real NCCL (`_update_sharded`) or, in single-GPU tests, an emulation that
        steps several ranks in lock-step."""
        sh = self.shard
        self._update_serial = getattr(self, "_update_serial", 0) + 1
        T, n_loc = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        assert n_loc == sh.n_local, "rollouts hold %d environments, the shard declares %d" % (n_loc, sh.n_local)
        batch_global = T * sh.n_total
        assert batch_global >= self.num_mini_batch
        mbs_g = batch_global // self.num_mini_batch
        n_mb = batch_global // mbs_g
        cap = getattr(self, "_cap", None) or _dist.default_capacity(mbs_g, sh)

        moments = _adv_moments(rt, rollouts)
        yield moments
        adv = _adv_apply(rt, rollouts, moments)
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        grads = rt.arena.bufs["G"]
        launches0 = _native.launch_count()
        for _ in range(self.ppo_epoch):
            perm = torch.randperm(batch_global).numpy()            # identical on every rank (same CPU seed)
            idx_np, counts = _dist.shard_epoch(perm, mbs_g, n_mb, sh, cap)
            if idx_np.shape[1] > cap:
                cap = (idx_np.shape[1] + 3) // 4 * 4
                idx_np, counts = _dist.shard_epoch(perm, mbs_g, n_mb, sh, cap)
            self._cap = cap
            prog, opt = self._program(rt, rollouts, cap, adv, accum, inv_B=1.0 / mbs_g)
            idx = torch.from_numpy(idx_np).to(rt.dev)
            for k in range(n_mb):
                prog.loss_desc.n_valid = int(counts[k])
                if getattr(prog, "fused_exchange", False):
                    self.optimizer.configure(prog.optim_desc)
                    prog.run(idx[k])
                else:
                    prog.run(idx[k])
                    yield grads
                    self.optimizer.configure(opt.optim_desc)
                    opt.run()
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
        self.last_launches = _native.launch_count() - launches0
        yield accum
        num_updates = self.ppo_epoch * self.num_mini_batch
        sums = accum[:3].cpu()
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)
