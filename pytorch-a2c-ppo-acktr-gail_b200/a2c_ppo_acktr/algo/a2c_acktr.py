"""A2C and ACKTR updates executed on the sm_100a engine.

Drop-in for `a2c_ppo_acktr.algo.A2C_ACKTR` (reference algo/a2c_acktr.py:9-80): one full-batch
forward / loss / backward over all T*N rows, then either gradient clipping + RMSprop (A2C) or the
Fisher-sampling backward + KFAC step (ACKTR, algo/kfac.py).
"""
import math

import torch

from .. import _dist, _native
from .kfac import KFACOptimizer
from .optim import FlatRMSprop


class A2C_ACKTR():
    def __init__(self,
                 actor_critic,
                 value_loss_coef,
                 entropy_coef,
                 lr=None,
                 eps=None,
                 alpha=None,
                 max_grad_norm=None,
                 acktr=False):
        self.actor_critic = actor_critic
        self.acktr = acktr
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        if acktr:
            self.optimizer = KFACOptimizer(actor_critic)
        else:
            self.optimizer = FlatRMSprop(actor_critic, lr, eps=eps, alpha=alpha)
        self._prog_key = None
        self._prog = None
        self.last_launches = 0
        self.shard = None

    def set_distributed(self, world, rank, env_range, n_total):
        """Environment-sharded update.  ACKTR: per-update factor increments and the
        gradient are all-reduced, every rank then performs the identical KFAC step.  A2C: loss terms are local
        sums over the GLOBAL batch size, the flat gradient is all-reduced before the (redundant, identical)
        clip + RMSprop step."""
        self.shard = _dist.ShardInfo(world, rank, env_range[0], env_range[1], n_total) if world > 1 else None
        self._prog_key = None
        if world > 1 and hasattr(self.actor_critic, "set_sampling_shard"):
            self.actor_critic.set_sampling_shard(env_range[0])      # exploration noise differs between ranks

    def update(self, rollouts):
        policy = self.actor_critic
        rt = policy.runtime()
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        B = T * N
        if getattr(rollouts, "wait_staged", None) is not None:
            rollouts.wait_staged()          # frames registered with stage_host(): upload them before the pass
        if self.acktr:
            return self._update_acktr(rt, rollouts)
        gen = self.a2c_steps(rt, rollouts)
        try:
            while True:
                _dist.all_reduce_sum(next(gen))
        except StopIteration as done:
            return done.value

    def a2c_steps(self, rt, rollouts):
        """reference a2c_acktr.py:38-80 with acktr=False.  Generator: yields every tensor that must be
        sum-all-reduced when the rollout is sharded (nothing is yielded on a single device)."""
        from .. import _engine
        policy = self.actor_critic
        sh = self.shard
        world = sh.world if sh is not None else 1
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        B = T * N
        assert sh is None or N == sh.n_local, "rollouts hold %d environments, the shard declares %d" % (N, sh.n_local)
        key = (id(rt), rollouts.obs.data_ptr(), rollouts.actions.data_ptr(), rollouts.returns.data_ptr(), B,
               rollouts.masks.data_ptr(), rollouts.recurrent_hidden_states.data_ptr(),
               self.value_loss_coef, self.entropy_coef, self.max_grad_norm, world)
        if key != self._prog_key:
            hyper = dict(value_loss_coef=self.value_loss_coef, entropy_coef=self.entropy_coef)
            if sh is not None:
                hyper["inv_B"] = 1.0 / (B * world)
            prog = rt.train_program(rollouts, B, 1, hyper, use_idx=False,
                                    seq=(T, N) if policy.is_recurrent else None)
            self.optimizer._ensure_state()
            opt = prog if sh is None else _engine.Program(rt.dev)     # sharded: the all-reduce sits in between
            opt.optim_desc = rt.add_optimizer_ops(opt, 1, self.optimizer.state1, None, self.max_grad_norm)
            prog.finalize()
            opt.finalize()
            self._prog, self._opt_prog, self._prog_key = prog, opt, key
        prog, opt = self._prog, self._opt_prog
        launches0 = _native.launch_count()
        if getattr(rollouts, "frame_ring", False):
            rt.stage_ring_batch(rollouts)            # the program reads materialised stacks at a fixed address
        self.optimizer.configure(opt.optim_desc)
        prog.run()
        if sh is not None:
            yield rt.arena.bufs["G"]
            opt.run()
            yield prog.loss_out
        self.last_launches = _native.launch_count() - launches0
        out = prog.loss_out[:3].cpu()
        return out[0].item(), out[1].item(), out[2].item()

    def _update_acktr(self, rt, rollouts):
        gen = self.acktr_steps(rt, rollouts)
        try:
            while True:
                t = next(gen)
                if self.shard is not None:
                    _dist.all_reduce_sum(t)
        except StopIteration as done:
            return done.value

    def acktr_steps(self, rt, rollouts):
        """reference a2c_acktr.py:38-80 with acktr=True.  Generator: yields every tensor that must be
        sum-all-reduced when the rollout is sharded (nothing is yielded on a single device)."""
        opt = self.optimizer
        sh = self.shard
        world = sh.world if sh is not None else 1
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        B = T * N
        lib, st = _native.lib(), _native.stream_ptr(rt.dev)
        obs_flat = rollouts.obs[:-1].view(B, -1)
        obs_f32 = rt.arena.ensure("obs_f32", obs_flat.numel())
        _native.check(lib.a2cb_frames_to_f32(obs_f32.data_ptr(), _native.dptr(obs_flat, name="obs"), obs_flat.numel(),
                                             int(obs_flat.dtype == torch.uint8), 255.0 if rt.spec.cnn else 1.0, st),
                      "a2cb_frames_to_f32")
        # value noise of the Fisher sample: drawn on the CPU from the global generator, as the
        # reference does (a2c_acktr.py:58-60), then uploaded
        noise = rt.arena.ensure("kfac_noise", B)
        use_fisher = opt.steps % opt.Ts == 0
        if use_fisher:
            if sh is None:
                noise.copy_(torch.randn(T, N, 1).view(-1), non_blocking=False)
            else:       # drawn for the GLOBAL batch (same seed on every rank), this shard's columns kept
                full = torch.randn(T, sh.n_total, 1)
                noise.copy_(full[:, sh.lo:sh.hi].contiguous().view(-1), non_blocking=False)
        key = (id(rt), rollouts.obs.data_ptr(), rollouts.actions.data_ptr(), rollouts.returns.data_ptr(), B,
               self.value_loss_coef, self.entropy_coef, world)
        if key != self._prog_key:
            hyper = dict(value_loss_coef=self.value_loss_coef, entropy_coef=self.entropy_coef)
            self._prog = rt.acktr_programs(rollouts, hyper, noise, inv_B=1.0 / (B * world))
            self._prog_key = key
        progs = self._prog
        opt._build(rt, rollouts, world)
        launches0 = _native.launch_count()
        progs["fwd"].run()
        opt.accumulate_a(rt)
        if use_fisher:
            opt.acc_stats = True
            progs["fisher"].run()
            opt.accumulate_g(rt)
            opt.acc_stats = False
        if sh is not None and use_fisher:
            # ~15.6 MB of factor statistics per update (CNN).  Only on steps that folded new statistics: each rank
            # then holds decay / world * m + (1 - decay) * its increment, so the SUM over ranks is the new average;
            # on any other step (Ts > 1) nothing was rescaled and a sum would multiply every factor by `world`.
            for fname, f in opt.factors.items():
                if fname != "one":
                    yield f["m"]
        progs["main"].run()
        if sh is not None:
            yield rt.arena.bufs["G"]
        opt.step(rt)
        self.last_launches = _native.launch_count() - launches0
        lo = progs["loss_out"]
        if sh is not None:
            yield lo
        out = lo[:3].cpu()
        return out[0].item(), out[1].item(), out[2].item()
