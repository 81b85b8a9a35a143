"""A2C update (and, in a later round, ACKTR) executed on the sm_100a engine.

Drop-in for `a2c_ppo_acktr.algo.A2C_ACKTR` (reference algo/a2c_acktr.py:9-80): one
full-batch forward / loss / backward over all T*N rows, gradient clipping and RMSprop.
"""
import torch

from .. import _native
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
            raise _native.NativeError("ACKTR (KFAC) is not built yet in this round")
        self.optimizer = FlatRMSprop(actor_critic, lr, eps=eps, alpha=alpha)
        self._prog_key = None
        self._prog = None
        self.last_launches = 0

    def update(self, rollouts):
        policy = self.actor_critic
        rt = policy.runtime()
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        B = T * N
        key = (id(rt), rollouts.obs.data_ptr(), rollouts.actions.data_ptr(), rollouts.returns.data_ptr(), B,
               rollouts.masks.data_ptr(), rollouts.recurrent_hidden_states.data_ptr(),
               self.value_loss_coef, self.entropy_coef, self.max_grad_norm)
        if key != self._prog_key:
            hyper = dict(value_loss_coef=self.value_loss_coef, entropy_coef=self.entropy_coef)
            prog = rt.train_program(rollouts, B, 1, hyper, use_idx=False,
                                    seq=(T, N) if policy.is_recurrent else None)
            self.optimizer._ensure_state()
            rt.add_optimizer_ops(prog, 1, self.optimizer.state1, None, self.max_grad_norm)
            prog.finalize()
            self._prog, self._prog_key = prog, key
        prog = self._prog
        launches0 = _native.launch_count()
        self.optimizer.configure(prog.optim_desc)
        prog.run()
        self.last_launches = _native.launch_count() - launches0
        out = prog.loss_out[:3].cpu()
        return out[0].item(), out[1].item(), out[2].item()
