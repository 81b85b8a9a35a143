"""GAIL discriminator and reward pass on the sm_100a engine.

Drop-in for `a2c_ppo_acktr.algo.gail.Discriminator` (reference algo/gail.py:11-110): same constructor
`(input_dim, hidden_dim, device)`, `update(expert_loader, rollouts, obsfilt=None) -> float`,
`predict_reward(state, action, gamma, masks, update_rms=True) -> [N, 1]`, a `trunk` whose `state_dict()`
keys and initialisation stream are the reference's, `returns` and `ret_rms`.  One minibatch step of `update`
is ONE kernel (three forward passes, both BCE terms, the gradient penalty with its double-backward terms in
closed form, all parameter gradients -- csrc/gail.cu) + a deterministic fold + the flat Adam kernel; the loss is
read back once per `update` instead of once per step.  `predict_rewards(rollouts, gamma)` (extension) replaces
the per-step Python loop of reference main.py:152-155 by two launches over the whole [T, N] rollout.

RNG consumption equals the reference's: `zip(expert_loader, policy batches)` starts the loader's iterator first,
the minibatch permutation (`feed_forward_generator(None, mini_batch_size=...)`, one `randperm(T * N)`) is drawn at
the first policy batch, and `alpha = torch.rand(B, 1)` comes from the global CPU generator inside every step.
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _engine, _native          # noqa: F401  (_engine registers the optimiser / fold entry points)

c_i32, c_f32, c_vp = ctypes.c_int32, ctypes.c_float, ctypes.c_void_p


class GailDesc(ctypes.Structure):
    """Mirror of a2cb_gail_t."""
    _fields_ = [("P", c_vp), ("slabs", c_vp), ("pol_state", c_vp), ("pol_action", c_vp), ("idx", c_vp),
                ("exp_state", c_vp), ("exp_action", c_vp), ("alpha", c_vp), ("raw_out", c_vp),
                ("B", c_i32), ("Ds", c_i32), ("Da", c_i32), ("H", c_i32), ("lambda_", c_f32), ("pad_", c_i32)]


_native.register({
    "a2cb_gail_disc": (c_i32, [c_i32, ctypes.POINTER(GailDesc), c_vp]),
    "a2cb_gail_slab_floats": (c_i32, [c_i32, c_i32]),
    "a2cb_gail_ctas": (c_i32, [c_i32]),
    "a2cb_gail_normalize": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_i32, c_vp]),
})

_KEYS = ("0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias")


class _RunningStats(object):
    """`ret_rms` of the reference (stable_baselines3 RunningMeanStd, shape ()): device-resident double
    {mean, var, count}, read back on attribute access."""

    def __init__(self, device):
        self.state = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device=device)

    @property
    def mean(self):
        return np.array([self.state[0].item()])

    @property
    def var(self):
        return np.array([self.state[1].item()])

    @property
    def count(self):
        return self.state[2].item()


class Discriminator(nn.Module):
    def __init__(self, input_dim, hidden_dim, device):
        super(Discriminator, self).__init__()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _native.NativeError("the discriminator lives on %s: pass a CUDA device (no CPU fallback)" % self.device)
        _native.lib()
        # same construction order / initialisation stream as the reference (three default-initialised nn.Linear)
        self.trunk = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.Tanh(),
            nn.Linear(hidden_dim, hidden_dim), nn.Tanh(),
            nn.Linear(hidden_dim, 1))
        self.input_dim, self.hidden_dim = int(input_dim), int(hidden_dim)
        self._flatten()
        self.trunk.train()
        # torch.optim.Adam defaults (reference gail.py:26): lr 1e-3, betas (0.9, 0.999), eps 1e-8
        self.optimizer = _FlatAdam(self)
        self.returns = None
        self.ret_rms = _RunningStats(self.device)
        self._returns_init = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._returns_buf = None

    # -- flat parameters ------------------------------------------------------------------------------------
    def _flatten(self):
        params = dict(self.trunk.named_parameters())
        n = sum(params[k].numel() for k in _KEYS)
        flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        grad = torch.zeros_like(flat)
        off = 0
        for k in _KEYS:
            p = params[k]
            m = p.numel()
            flat[off:off + m].copy_(p.data.reshape(-1))
            p.data = flat[off:off + m].view(p.shape)
            p.grad = grad[off:off + m].view(p.shape)
            off += m
        self._flat, self._flat_grad = flat, grad

    def _is_flat(self):
        off = 0
        params = dict(self.trunk.named_parameters())
        for k in _KEYS:
            p = params[k]
            if p.data_ptr() != self._flat.data_ptr() + 4 * off or p.grad is None:
                return False
            off += p.numel()
        return True

    def _apply(self, fn, *a, **k):
        out = super(Discriminator, self)._apply(fn, *a, **k)
        self._flatten()
        return out

    # -- reference gail.py:57-96 ------------------------------------------------------------------------------
    def update(self, expert_loader, rollouts, obsfilt=None):
        self.train()
        if not self._is_flat():
            self._flatten()
        lib, dev = _native.lib(), self.device
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        obs = rollouts.obs[:-1].reshape(T * N, -1)
        act = rollouts.actions.reshape(T * N, -1).to(torch.float32)
        if obs.dtype != torch.float32 or not obs.is_cuda:
            raise _native.NativeError("GAIL reads float32 CUDA observations")
        obs, act = obs.contiguous(), act.contiguous()
        Ds, Da = obs.shape[1], act.shape[1]
        if Ds + Da != self.input_dim:
            raise ValueError("discriminator input is %d wide, state + action is %d" % (self.input_dim, Ds + Da))
        B = expert_loader.batch_size
        n_cta = int(lib.a2cb_gail_ctas(B))
        span = int(lib.a2cb_gail_slab_floats(self.input_dim, self.hidden_dim))
        slabs = torch.empty(n_cta * span, device=dev)
        folded = torch.empty(span, device=dev)
        loss_sum = torch.zeros(1, device=dev)
        n_param = self._flat.numel()

        def policy_batches():
            # feed_forward_generator(None, mini_batch_size=B): ONE randperm drawn when iteration starts, consecutive
            # chunks, remainder dropped (reference storage.py:122-125) -- only the row lists are needed here, the kernel
            # gathers the two fields it reads
            perm = torch.randperm(T * N)
            for k in range((T * N) // B):
                yield perm[k * B:(k + 1) * B]

        n = 0
        keep = []
        for expert_batch, idx in zip(expert_loader, policy_batches()):
            expert_state, expert_action = expert_batch
            if obsfilt is not None:
                expert_state = torch.FloatTensor(obsfilt(expert_state.numpy(), update=False))
            if expert_state.shape[0] != B:
                raise ValueError("expert batch of %d rows, policy batch of %d (the reference's mixup needs equal sizes)"
                                 % (expert_state.shape[0], B))
            es = expert_state.to(dev, torch.float32).contiguous()
            ea = expert_action.to(dev, torch.float32).contiguous()
            alpha = torch.rand(B, 1).to(dev).view(-1).contiguous()          # reference gail.py:37
            idx_d = idx.to(dev)
            d = GailDesc()
            d.P, d.slabs = self._flat.data_ptr(), slabs.data_ptr()
            d.pol_state, d.pol_action, d.idx = obs.data_ptr(), act.data_ptr(), idx_d.data_ptr()
            d.exp_state, d.exp_action, d.alpha = es.data_ptr(), ea.data_ptr(), alpha.data_ptr()
            d.B, d.Ds, d.Da, d.H, d.lambda_ = B, Ds, Da, self.hidden_dim, 10.0
            st = _native.stream_ptr(dev)
            with torch.cuda.device(dev):
                _native.check(lib.a2cb_gail_disc(0, ctypes.byref(d), st), "a2cb_gail_disc")
                _native.check(lib.a2cb_reduce_splits(folded.data_ptr(), slabs.data_ptr(), span, span, n_cta, 0, None, 0, 0,
                                                     1.0, st), "a2cb_reduce_splits")
            self._flat_grad.copy_(folded[:n_param])
            loss_sum += folded[n_param] + folded[n_param + 1]
            self.optimizer.step()
            keep.append((es, ea, alpha, idx_d))                             # alive until the stream has consumed them
            n += 1
        if n == 0:
            raise ZeroDivisionError("division by zero")                     # the reference's `loss / n`
        return loss_sum.item() / n

    # -- reference gail.py:98-110 ------------------------------------------------------------------------------
    def _raw_rewards(self, state, action):
        lib, dev = _native.lib(), self.device
        if not state.is_cuda:
            raise _native.NativeError("GAIL reads CUDA tensors; there is no CPU fallback")
        state = state.to(torch.float32).contiguous()
        action = action.to(torch.float32).contiguous().view(state.shape[0], -1)
        B = state.shape[0]
        raw = torch.empty(B, device=dev)
        d = GailDesc()
        d.P, d.pol_state, d.pol_action, d.raw_out = self._flat.data_ptr(), state.data_ptr(), action.data_ptr(), raw.data_ptr()
        d.B, d.Ds, d.Da, d.H = B, state.shape[1], action.shape[1], self.hidden_dim
        if d.Ds + d.Da != self.input_dim:
            raise ValueError("discriminator input is %d wide, state + action is %d" % (self.input_dim, d.Ds + d.Da))
        with torch.cuda.device(dev):
            _native.check(lib.a2cb_gail_disc(1, ctypes.byref(d), _native.stream_ptr(dev)), "a2cb_gail_disc")
        self._keep = (state, action)
        return raw

    def _normalize(self, raw, masks, T, N, gamma, update_rms):
        lib, dev = _native.lib(), self.device
        if self._returns_buf is None or self._returns_buf.numel() != N:
            self._returns_buf = torch.zeros(N, device=dev)
            self._returns_init.zero_()
        out = torch.empty(T * N, device=dev)
        m = masks.to(torch.float32).contiguous().view(-1)
        with torch.cuda.device(dev):
            _native.check(lib.a2cb_gail_normalize(out.data_ptr(), raw.data_ptr(), m.data_ptr(), T, N, float(gamma),
                                                  self._returns_buf.data_ptr(), self._returns_init.data_ptr(),
                                                  self.ret_rms.state.data_ptr(), int(bool(update_rms)),
                                                  _native.stream_ptr(dev)), "a2cb_gail_normalize")
        self.returns = self._returns_buf.view(N, 1)
        return out

    def predict_reward(self, state, action, gamma, masks, update_rms=True):
        with torch.no_grad():
            self.eval()
            if not self._is_flat():
                self._flatten()
            N = state.shape[0]
            raw = self._raw_rewards(state, action)
            return self._normalize(raw, masks, 1, N, gamma, update_rms).view(N, 1)

    def predict_rewards(self, rollouts, gamma, update_rms=True):
        """The loop of reference main.py:152-155 over a whole rollout: `rollouts.rewards[step] = predict_reward(obs[step],
        actions[step], gamma, masks[step])` for every step, as two launches (all T * N raw rewards, then the T sequential
        running-statistics steps).  Writes `rollouts.rewards` in place and returns it."""
        with torch.no_grad():
            self.eval()
            if not self._is_flat():
                self._flatten()
            T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
            raw = self._raw_rewards(rollouts.obs[:-1].reshape(T * N, -1), rollouts.actions.reshape(T * N, -1))
            out = self._normalize(raw, rollouts.masks[:-1], T, N, gamma, update_rms)
            rollouts.rewards.copy_(out.view(T, N, 1))
            return rollouts.rewards


class _FlatAdam(object):
    """torch.optim.Adam(self.trunk.parameters()) with its defaults, as ONE launch on the flat buffer (csrc/optim.cu)."""

    def __init__(self, disc, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.disc = disc
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, params=list(disc.trunk.parameters()))]
        self.step_count = 0
        self.state1 = self.state2 = None
        self._norm = None

    def zero_grad(self, set_to_none=False):
        self.disc._flat_grad.zero_()

    def step(self):
        d = self.disc
        flat, grad = d._flat, d._flat_grad
        if self.state1 is None or self.state1.data_ptr() == 0 or self.state1.numel() != flat.numel() or self.state1.device != flat.device:
            self.state1, self.state2 = torch.zeros_like(flat), torch.zeros_like(flat)
            self._norm = torch.zeros(4, device=flat.device)
        self.step_count += 1
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        t = self.step_count
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        with torch.cuda.device(flat.device):
            rc = _native.lib().a2cb_optim_step(0, flat.data_ptr(), grad.data_ptr(), self.state1.data_ptr(), self.state2.data_ptr(),
                                               flat.numel(), self._norm.data_ptr(), g["lr"] / bc1, g["eps"], b1, b2, 1 - b1, 1 - b2,
                                               math.sqrt(bc2), 0.0, int(t == 1), _native.stream_ptr(flat.device))
        _native.check(rc, "a2cb_optim_step")
