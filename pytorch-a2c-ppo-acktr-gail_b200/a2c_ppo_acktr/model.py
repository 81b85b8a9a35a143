"""Actor-critic policy with the reference's public surface, executed by sm_100a kernels.

Drop-in for `a2c_ppo_acktr.model.Policy` (reference model.py:15-79): same constructor,
`act / get_value / evaluate_actions`, `is_recurrent`, `recurrent_hidden_state_size`, and
a `state_dict()` whose keys and shapes equal the reference's, so checkpoints move both
ways.  Initialisation consumes the global RNG exactly like the reference (same layer
construction order, orthogonal init with the same gains, zero biases).

All parameters are views into ONE flat fp32 buffer (and their .grad into one flat gradient
buffer) so that the fused update needs a single norm / optimiser launch and, when sharded,
a single NCCL all-reduce.  The dense math runs on the affine-indexed GEMM engine
(csrc/gemm.cu); log-prob / entropy / losses on the fused epilogue (csrc/loss.cu).
There is no CPU path: calling the policy with CPU tensors raises.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _engine, _native
from .distributions import Categorical, DiagGaussian, _holder
from .utils import init


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class _Spec(object):
    def __init__(self, obs_shape, action_kind, num_actions, recurrent, hidden):
        self.obs_shape = tuple(obs_shape)
        self.cnn = len(self.obs_shape) == 3
        self.action_kind, self.num_actions = action_kind, num_actions
        self.recurrent, self.hidden = bool(recurrent), hidden


class NNBase(nn.Module):
    def __init__(self, recurrent, recurrent_input_size, hidden_size):
        super(NNBase, self).__init__()
        self._hidden_size = hidden_size
        self._recurrent = recurrent
        if recurrent:
            # reference model.py:89-95: GRU drawn first, orthogonal weights, zero biases
            gru = nn.GRU(recurrent_input_size, hidden_size)
            for name, param in gru.named_parameters():
                if 'bias' in name:
                    nn.init.constant_(param, 0)
                elif 'weight' in name:
                    nn.init.orthogonal_(param)
            self.gru = nn.Module()
            for name, param in gru.named_parameters():
                setattr(self.gru, name, nn.Parameter(param.data.clone()))

    @property
    def is_recurrent(self):
        return self._recurrent

    @property
    def recurrent_hidden_state_size(self):
        return self._hidden_size if self._recurrent else 1

    @property
    def output_size(self):
        return self._hidden_size


def _sequential(named_layers):
    seq = nn.Module()
    for name, layer in named_layers:
        seq.add_module(name, _holder(layer))
    return seq


class CNNBase(NNBase):
    """Nature-DQN trunk (reference model.py:169-195): conv 8x8/4 -> 4x4/2 -> 3x3/1 -> fc 512."""

    def __init__(self, num_inputs, recurrent=False, hidden_size=512):
        super(CNNBase, self).__init__(recurrent, hidden_size, hidden_size)
        g = nn.init.calculate_gain('relu')
        zero = lambda x: nn.init.constant_(x, 0)
        self.main = _sequential([
            ('0', init(nn.Conv2d(num_inputs, 32, 8, stride=4), nn.init.orthogonal_, zero, g)),
            ('2', init(nn.Conv2d(32, 64, 4, stride=2), nn.init.orthogonal_, zero, g)),
            ('4', init(nn.Conv2d(64, 32, 3, stride=1), nn.init.orthogonal_, zero, g)),
            ('7', init(nn.Linear(32 * 7 * 7, hidden_size), nn.init.orthogonal_, zero, g)),
        ])
        self.critic_linear = _holder(init(nn.Linear(hidden_size, 1), nn.init.orthogonal_, zero))
        self.train()


class MLPBase(NNBase):
    """Separate tanh actor / critic trunks (reference model.py:198-229)."""

    def __init__(self, num_inputs, recurrent=False, hidden_size=64):
        super(MLPBase, self).__init__(recurrent, num_inputs, hidden_size)
        if recurrent:
            num_inputs = hidden_size
        g = np.sqrt(2)
        zero = lambda x: nn.init.constant_(x, 0)
        mk = lambda i, o: init(nn.Linear(i, o), nn.init.orthogonal_, zero, g)
        self.actor = _sequential([('0', mk(num_inputs, hidden_size)), ('2', mk(hidden_size, hidden_size))])
        self.critic = _sequential([('0', mk(num_inputs, hidden_size)), ('2', mk(hidden_size, hidden_size))])
        self.critic_linear = _holder(mk(hidden_size, 1))
        self.train()


class Policy(nn.Module):
    def __init__(self, obs_shape, action_space, base=None, base_kwargs=None):
        super(Policy, self).__init__()
        if base_kwargs is None:
            base_kwargs = {}
        if base is None:
            if len(obs_shape) == 3:
                base = CNNBase
            elif len(obs_shape) == 1:
                base = MLPBase
            else:
                raise NotImplementedError
        self.base = base(obs_shape[0], **base_kwargs)

        space = action_space.__class__.__name__
        if space == "Discrete":
            self.dist = Categorical(self.base.output_size, action_space.n)
            kind, na = "discrete", action_space.n
        elif space == "Box":
            self.dist = DiagGaussian(self.base.output_size, action_space.shape[0])
            kind, na = "box", action_space.shape[0]
        else:
            # MultiBinary/Bernoulli: out of scope (broken upstream, distributions.py:50)
            raise NotImplementedError
        self._spec = _Spec(obs_shape, kind, na, self.base.is_recurrent, self.base.output_size)
        self._layout = _engine.ParamLayout(self._spec)
        self._flat = None
        self._flat_grad = None
        self._runtime = None
        self._flatten()

    # -- flat parameter buffer ------------------------------------------------
    def _named_by_key(self):
        return dict(self.named_parameters())

    def _flatten(self):
        """(Re)creates the flat buffers on the parameters' current device and re-points every
        parameter (and gradient) at its slice; Parameter identity is preserved."""
        params = self._named_by_key()
        dev = next(iter(params.values())).device
        flat = torch.zeros(self._layout.total, dtype=torch.float32, device=dev)
        grad = torch.zeros_like(flat)
        for key, (off, shape) in self._layout.entries.items():
            p = params[key]
            n = p.numel()
            assert tuple(p.shape) == tuple(shape), (key, p.shape, shape)
            flat[off:off + n].copy_(p.data.reshape(-1))
            if p.grad is not None:
                grad[off:off + n].copy_(p.grad.reshape(-1))
            p.data = flat[off:off + n].view(shape)
            p.grad = grad[off:off + n].view(shape)
        assert len(params) == len(self._layout.entries)
        self._flat, self._flat_grad = flat, grad
        self._runtime = None

    def _is_flat(self):
        if self._flat is None:
            return False
        base = self._flat.data_ptr()
        gbase = self._flat_grad.data_ptr()
        params = self._named_by_key()
        for key, (off, _) in self._layout.entries.items():
            p = params[key]
            if p.data_ptr() != base + 4 * off or p.grad is None or p.grad.data_ptr() != gbase + 4 * off:
                return False
        return True

    def _apply(self, fn, *a, **k):
        out = super(Policy, self)._apply(fn, *a, **k)
        self._flatten()
        return out

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_flat"] = st["_flat_grad"] = st["_runtime"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._flatten()

    def runtime(self):
        if not self._is_flat():
            self._flatten()
        if self._runtime is None:
            self._runtime = _engine.PolicyRuntime(self._spec, self._layout, self._flat, self._flat_grad)
        return self._runtime

    # -- reference surface -----------------------------------------------------
    @property
    def is_recurrent(self):
        return self.base.is_recurrent

    @property
    def recurrent_hidden_state_size(self):
        return self.base.recurrent_hidden_state_size

    def forward(self, inputs, rnn_hxs, masks):
        raise NotImplementedError

    def act(self, inputs, rnn_hxs, masks, deterministic=False):
        """-> (value [N,1], action [N,1] int64 | [N,A] fp32, action_log_probs [N,1], rnn_hxs)."""
        rt = self.runtime()
        with torch.no_grad():
            out = rt.forward(inputs, rnn_hxs, masks)
            z = out["z"]
            if self._spec.action_kind == "discrete":
                logits = z[:, 1:]
                if deterministic:
                    action = logits.argmax(dim=-1, keepdim=True)
                else:
                    action = torch.multinomial(torch.softmax(logits, dim=-1), 1, True)
            else:
                mean = z[:, 1:]
                if deterministic:
                    action = mean.clone()
                else:
                    std = self.dist.logstd._bias.detach().t().exp().expand_as(mean)
                    action = torch.normal(mean, std)
            stats = rt.head_stats(z.shape[0], action)
        return z[:, 0:1].clone(), action, stats["logp"].view(-1, 1).clone(), out["hxs"]

    def get_value(self, inputs, rnn_hxs, masks):
        with torch.no_grad():
            return self.runtime().forward(inputs, rnn_hxs, masks)["z"][:, 0:1].clone()

    def evaluate_actions(self, inputs, rnn_hxs, masks, action):
        """-> (value [B,1], action_log_probs [B,1], dist_entropy (scalar mean), rnn_hxs).
        Differentiable w.r.t. the parameters when grad mode is on."""
        rt = self.runtime()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            params = [p for _, p in sorted(self._named_by_key().items())]
            keys = sorted(self._named_by_key().keys())
            value, logp, ent, hxs = _EvaluateActions.apply(self, keys, inputs, rnn_hxs, masks, action, *params)
            return value, logp, ent, hxs
        with torch.no_grad():
            out = rt.forward(inputs, rnn_hxs, masks)
            stats = rt.head_stats(out["z"].shape[0], action)
        return out["z"][:, 0:1].clone(), stats["logp"].view(-1, 1).clone(), stats["entropy"].clone(), out["hxs"]


class _EvaluateActions(torch.autograd.Function):
    """Autograd bridge: forward and backward both run on the native engine."""

    @staticmethod
    def forward(ctx, policy, keys, inputs, rnn_hxs, masks, action, *params):
        rt = policy.runtime()
        out = rt.forward(inputs, rnn_hxs, masks, keep_for_backward=True)
        B = out["z"].shape[0]
        stats = rt.head_stats(B, action)
        ctx.policy, ctx.keys, ctx.B, ctx.action = policy, keys, B, action
        ctx.token = rt.forward_token
        ctx.mark_non_differentiable(out["hxs"])
        return out["z"][:, 0:1].clone(), stats["logp"].view(-1, 1).clone(), stats["entropy"].clone(), out["hxs"]

    @staticmethod
    def backward(ctx, g_value, g_logp, g_ent, _g_hxs):
        policy = ctx.policy
        rt = policy.runtime()
        if rt.forward_token != ctx.token:
            raise RuntimeError("evaluate_actions: activations were overwritten by a later forward pass; "
                               "call backward() before evaluating another batch of the same size")
        B = ctx.B
        dev = policy._flat.device
        gv = (g_value if g_value is not None else torch.zeros(B, 1, device=dev)).reshape(-1).contiguous().float()
        gl = (g_logp if g_logp is not None else torch.zeros(B, 1, device=dev)).reshape(-1).contiguous().float()
        ge = (g_ent if g_ent is not None else torch.zeros((), device=dev)).reshape(1).contiguous().float()
        # The engine writes parameter gradients into the flat buffer that p.grad aliases; autograd
        # will ACCUMULATE what we return into p.grad, so hand it copies and put the buffer back.
        flat_grad = policy._flat_grad
        before = flat_grad.clone()
        grads = rt.backward_vjp(B, ctx.action, gv, gl, ge).clone()
        flat_grad.copy_(before)
        by_key = {}
        for key, (off, shape) in policy._layout.entries.items():
            n = int(math.prod(shape))
            by_key[key] = grads[off:off + n].view(shape)
        return (None, None, None, None, None, None) + tuple(by_key[k] for k in ctx.keys)
