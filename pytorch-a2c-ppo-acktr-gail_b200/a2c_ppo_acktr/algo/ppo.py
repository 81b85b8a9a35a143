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
import torch

from .. import _native
from .optim import FlatAdam


def _adv_normalize(rt, rollouts, allreduce=None):
    """(returns - value_preds - mean) / (unbiased std + 1e-5) over all T*N rows (ppo.py:35-37)."""
    T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
    n = T * N
    adv = rt.arena.ensure("adv", n)
    scratch = rt.arena.ensure("adv_scratch", 2 * 128 + 2, torch.float64)
    moments = rt.arena.ensure("adv_moments", 4, torch.float64)
    lib = _native.lib()
    st = _native.stream_ptr(rt.dev)
    with torch.cuda.device(rt.dev):
        _native.check(lib.a2cb_adv_moments(_native.dptr(rollouts.returns, torch.float32, "returns"),
                                           _native.dptr(rollouts.value_preds, torch.float32, "value_preds"),
                                           n, scratch.data_ptr(), moments.data_ptr(), st), "a2cb_adv_moments")
        if allreduce is not None:
            allreduce(moments)
        _native.check(lib.a2cb_adv_normalize(rollouts.returns.data_ptr(), rollouts.value_preds.data_ptr(), n,
                                             moments.data_ptr(), adv.data_ptr(), st), "a2cb_adv_normalize")
    return adv


class PPO():
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

    def _program(self, rt, rollouts, mbs, adv, accum):
        key = (id(rt), rollouts.obs.data_ptr(), rollouts.actions.data_ptr(), rollouts.returns.data_ptr(),
               rollouts.value_preds.data_ptr(), rollouts.action_log_probs.data_ptr(), adv.data_ptr(),
               accum.data_ptr(), mbs, self.clip_param, self.value_loss_coef, self.entropy_coef,
               self.use_clipped_value_loss, self.max_grad_norm)
        if key != self._prog_key:
            hyper = dict(clip_param=self.clip_param, value_loss_coef=self.value_loss_coef,
                         entropy_coef=self.entropy_coef, use_clipped_value_loss=self.use_clipped_value_loss)
            prog = rt.train_program(rollouts, mbs, 0, hyper, adv=adv, loss_accum=accum)
            self.optimizer._ensure_state()
            rt.add_optimizer_ops(prog, 0, self.optimizer.state1, self.optimizer.state2, self.max_grad_norm)
            prog.finalize()
            self._prog, self._prog_key = prog, key
        return self._prog

    def update(self, rollouts):
        policy = self.actor_critic
        rt = policy.runtime()
        if policy.is_recurrent:
            raise _native.NativeError("recurrent PPO is not built yet in this round")
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        batch_size = T * N
        assert batch_size >= self.num_mini_batch, (
            "PPO requires the number of processes ({}) "
            "* number of steps ({}) = {} "
            "to be greater than or equal to the number of PPO mini batches ({})."
            "".format(N, T, N * T, self.num_mini_batch))
        mbs = batch_size // self.num_mini_batch
        n_mb = batch_size // mbs

        adv = _adv_normalize(rt, rollouts)
        accum = rt.arena.ensure("loss_accum", 4)
        accum.zero_()
        prog = self._program(rt, rollouts, mbs, adv, accum)
        launches0 = _native.launch_count()
        for _ in range(self.ppo_epoch):
            perm = torch.randperm(batch_size).to(rt.dev)
            for k in range(n_mb):
                self.optimizer.configure(prog.optim_desc)
                prog.run(perm[k * mbs:(k + 1) * mbs])
                if self.step_hook is not None:
                    self.step_hook(self.optimizer.step_count)
        self.last_launches = _native.launch_count() - launches0
        num_updates = self.ppo_epoch * self.num_mini_batch
        sums = accum[:3].cpu()          # the update's only device->host synchronisation
        return (sums[0].item() / num_updates, sums[1].item() / num_updates, sums[2].item() / num_updates)
