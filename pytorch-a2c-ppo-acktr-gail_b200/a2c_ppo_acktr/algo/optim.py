"""Flat-buffer optimisers backed by csrc/optim.cu.

They keep the slice of the torch.optim surface the reference's callers touch:
`param_groups[i]['lr']` is writable and honoured at every step
(`utils.update_linear_schedule`, reference main.py:107-111), `zero_grad()`, `step()`,
`state_dict()`.  Arithmetic: torch 2.11 single-tensor Adam / RMSprop (see csrc/optim.cu).
"""
import ctypes
import math

import torch

from .. import _native


class _FlatOptimizer(object):
    KIND = None

    def __init__(self, policy, defaults):
        self.policy = policy
        self.param_groups = [dict(defaults, params=list(policy.parameters()))]
        self.step_count = 0
        self._state_dev = None
        self.state1 = self.state2 = None

    def _ensure_state(self):
        flat = self.policy.runtime().arena.bufs["P"]
        if self.state1 is None or self.state1.device != flat.device or self.state1.numel() != flat.numel():
            self.state1 = torch.zeros_like(flat)
            self.state2 = torch.zeros_like(flat) if self.KIND == 0 else None
        return flat

    def zero_grad(self, set_to_none=False):
        self.policy.runtime().arena.bufs["G"].zero_()

    def scalars(self):
        raise NotImplementedError

    def configure(self, desc):
        """Writes this step's scalars into an a2cb_optim_t (called once per optimiser step)."""
        self.step_count += 1
        for k, v in self.scalars().items():
            setattr(desc, k, v)

    def step(self, max_norm=None):
        """Standalone step on the policy's flat gradient buffer (the fused update path appends the
        same ops to its op list instead)."""
        rt = self.policy.runtime()
        flat = self._ensure_state()
        from .. import _engine
        prog = _engine.Program(flat.device)
        desc = rt.add_optimizer_ops(prog, self.KIND, self.state1, self.state2, max_norm)
        self.configure(desc)
        prog.run()

    def state_dict(self):
        return {"step": self.step_count, "state1": self.state1, "state2": self.state2,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.state1, self.state2 = sd["state1"], sd["state2"]
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


class FlatAdam(_FlatOptimizer):
    KIND = 0

    def __init__(self, policy, lr, eps, betas=(0.9, 0.999)):
        super(FlatAdam, self).__init__(policy, dict(lr=lr, eps=eps, betas=betas))

    def scalars(self):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        t = self.step_count
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        return dict(lr_eff=g["lr"] / bc1, eps=g["eps"], b1=b1, b2=b2, omb1=1 - b1, omb2=1 - b2,
                    bc2_sqrt=math.sqrt(bc2), first_step=int(t == 1))


class FlatRMSprop(_FlatOptimizer):
    KIND = 1

    def __init__(self, policy, lr, eps, alpha):
        super(FlatRMSprop, self).__init__(policy, dict(lr=lr, eps=eps, alpha=alpha))

    def scalars(self):
        g = self.param_groups[0]
        return dict(lr_eff=g["lr"], eps=g["eps"], b1=0.0, b2=g["alpha"], omb1=1.0, omb2=1 - g["alpha"],
                    bc2_sqrt=1.0, first_step=int(self.step_count == 1))
