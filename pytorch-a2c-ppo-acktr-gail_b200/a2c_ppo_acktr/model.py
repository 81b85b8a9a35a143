"""Actor-critic policy with the reference's public surface, executed by sm_100a kernels.

Drop-in for `a2c_ppo_acktr.model.Policy` (reference model.py:15-79): same constructor,
`act / get_value / evaluate_actions`, `is_recurrent`, `recurrent_hidden_state_size`, and
a `state_dict()` whose keys and shapes equal the reference's, so checkpoints move both
ways.  Initialisation consumes the global RNG exactly like the reference (same layer
construction order, orthogonal init with the same gains, zero biases).

All parameters are views into ONE flat fp32 buffer (and their .grad into one flat gradient
buffer) so that the fused update needs a single norm / optimiser launch and, when sharded,
a single NCCL all-reduce.  The CNN trunk and the GRU input product run on the TMA-fed tcgen05 engine
(csrc/tcx.cu), the GRU recurrence on csrc/gru_*.cu, MLP policies on csrc/mlp_fused.cu, everything else on
the affine-indexed GEMM engine (csrc/gemm*.cu); log-prob / entropy / losses on the fused epilogue (csrc/loss.cu).
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
        self._rng = None             # sampling stream of act(): see sampling_state()
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
        rng = st.get("_rng")
        if rng is not None:          # a checkpoint resumes the sampling stream where it stopped
            st["_rng"] = dict(rng, state=rng["state"].cpu())
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.__dict__.setdefault("_rng", None)
        self._flatten()

    def runtime(self):
        if not self._is_flat():
            self._flatten()
        if self._runtime is None:
            self._runtime = _engine.PolicyRuntime(self._spec, self._layout, self._flat, self._flat_grad)
            self._runtime.rng_provider = self.sampling_state
        return self._runtime

    # -- sampling stream of act() (csrc/actor.cu) -------------------------------
    def sampling_state(self):
        """Device int64 {seed, offset, ticket, row_base}: Philox key, launches drawn so far, the kernel's ticket
        counter and the GLOBAL index of this policy's first environment (`set_sampling_shard`).  The state belongs to
        the policy, not to its runtime: it survives `_flatten()`, `.to(device)` and pickling (`torch.save` of the
        policy, reference main.py:173-176).  The key is derived from `torch.initial_seed()`; a later
        `torch.manual_seed` is noticed here and restarts the stream with the new key."""
        dev = self._flat.device
        cpu_seed = torch.initial_seed()
        rng = self._rng
        if rng is None:
            rng = self._rng = dict(cpu_seed=cpu_seed, state=torch.tensor(
                [_engine.derive_sampling_seed(cpu_seed), 0, 0, 0], dtype=torch.int64, device=dev))
        if rng["state"].device != dev:
            rng["state"] = rng["state"].to(dev)
            self._runtime = None     # cached actor programs hold the old device address
        if rng["cpu_seed"] != cpu_seed:
            base = int(rng["state"][3].item())
            rng["state"].copy_(torch.tensor([_engine.derive_sampling_seed(cpu_seed), 0, 0, base], dtype=torch.int64))
            rng["cpu_seed"] = cpu_seed
        return rng["state"]

    def set_sampling_shard(self, first_env):
        """Environment-sharded actors: rank r passes the global index of its first environment, so that identically
        seeded ranks draw DIFFERENT exploration noise for different environments (counter = first_env + local row)."""
        self.sampling_state()[3] = int(first_env)

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
        """-> (value [N,1], action [N,1] int64 | [N,A] fp32, action_log_probs [N,1], rnn_hxs).
        One native program: no-grad forward + the sampling head (csrc/actor.cu).  Inputs are copied into the
        program's own buffers, so one cached program per batch size serves every step of a rollout."""
        rt = self.runtime()
        if not inputs.is_cuda:
            raise _native.NativeError("observations must be CUDA tensors; there is no CPU fallback")
        B, s = inputs.shape[0], self._spec
        if tuple(inputs.shape[1:]) != s.obs_shape:
            raise ValueError("expected observations of shape [B, %s], got %s" % (s.obs_shape, tuple(inputs.shape)))
        rt.obs_code(inputs)                                   # dtype check
        if not deterministic:
            self.sampling_state()                             # notices a re-seeded CPU generator
        if s.recurrent and rnn_hxs.shape[0] != B:
            return self._act_sequence(inputs, rnn_hxs, masks, deterministic)
        with torch.no_grad():
            plan = rt.act_plan(B, inputs.dtype)
            t = plan.tensors
            t["obs"].copy_(inputs.reshape(-1))
            if s.recurrent:
                t["hxs"].copy_(rnn_hxs.reshape(-1))
                t["masks"].copy_(masks.reshape(-1))
            plan.run(deterministic)
            rt.forward_token += 1
            hxs = t["hout"].view(B, s.hidden).clone() if s.recurrent else rnn_hxs
            action = t["action"].view(B, -1).clone()
            return t["value"].view(B, 1).clone(), action, t["logp"].view(B, 1).clone(), hxs

    def _act_sequence(self, inputs, rnn_hxs, masks, deterministic):
        """act() on a [T*N] batch of a recurrent policy (not used by the reference's actor loop): generic forward,
        then the sampling head on its outputs."""
        rt = self.runtime()
        with torch.no_grad():
            out = rt.forward(inputs, rnn_hxs, masks)
            z = out["z"]
            B = z.shape[0]
            action = torch.zeros(B, 1, dtype=torch.int64, device=z.device) if self._spec.action_kind == "discrete" \
                else torch.zeros(B, self._spec.num_actions, device=z.device)
            value, logp = torch.zeros(B, 1, device=z.device), torch.zeros(B, 1, device=z.device)
            rt.sample(z, action, value, logp, deterministic)
        return value, action, logp, out["hxs"]

    def act_into(self, rollouts, step=None, deterministic=False):
        """Fused actor step (extension; reference main.py:115-134 = act + the policy half of insert): reads
        `rollouts.obs[step]`, `.recurrent_hidden_states[step]`, `.masks[step]` IN PLACE and writes
        `.value_preds[step]`, `.actions[step]`, `.action_log_probs[step]` and `.recurrent_hidden_states[step + 1]`
        directly -- one native program, no staging copies, no temporaries.  Returns views of those slots in the
        order of act(); passing them on to `rollouts.insert(...)` is harmless (copies onto themselves)."""
        rt = self.runtime()
        s = self._spec
        step = rollouts.step if step is None else step
        wait = getattr(rollouts, "wait_staged", None)
        if wait is not None:
            wait()
        # frame ring: the slot's stack is assembled in a persistent staging tensor (stable address for the graph)
        obs = rollouts.ring_stack(step) if getattr(rollouts, "frame_ring", False) else rollouts.obs[step]
        hxs, masks = rollouts.recurrent_hidden_states[step], rollouts.masks[step]
        value, action = rollouts.value_preds[step], rollouts.actions[step]
        logp, hout = rollouts.action_log_probs[step], rollouts.recurrent_hidden_states[step + 1]
        B = obs.shape[0]
        want_act = torch.int64 if s.action_kind == "discrete" else torch.float32
        ok = (obs.is_cuda and tuple(obs.shape[1:]) == s.obs_shape and action.dtype == want_act
              and all(x.is_contiguous() for x in (obs, hxs, masks, value, action, logp, hout))
              and obs.data_ptr() % 16 == 0 and (not s.recurrent or (hxs.data_ptr() % 16 == 0 and hout.data_ptr() % 16 == 0))
              and obs.numel() * obs.element_size() <= rt.STAGE_BYTES)
        if not ok:                                            # odd layouts: the copying path
            v, a, lp, h = self.act(obs, hxs, masks, deterministic)
            for dst, src in ((value, v), (action, a), (logp, lp), (hout, h)):
                dst.copy_(src)
            return value, action, logp, hout
        rt.obs_code(obs)
        if not deterministic:
            self.sampling_state()
        with torch.no_grad():
            plan = rt.act_plan(B, obs.dtype)
            ptrs = {"obs": obs.data_ptr(), "value": value.data_ptr(), "action": action.data_ptr(), "logp": logp.data_ptr()}
            if s.recurrent:
                ptrs.update(hxs=hxs.data_ptr(), masks=masks.data_ptr(), hout=hout.data_ptr())
            def launch():
                plan.run(deterministic, ptrs)
                if not s.recurrent:
                    hout.copy_(hxs)
            self._act_graphed(plan, (bool(deterministic),) + tuple(sorted(ptrs.items())), launch)
            rt.forward_token += 1
        return value, action, logp, hout

    use_cuda_graph = True      # act_into: one CUDA graph per (rollout slot, mode), captured on its second use
    MAX_ACT_GRAPHS = 512       # per batch size: a rollout has num_steps slots (2048 for the MuJoCo recipe runs uncaptured
                               # beyond this); the least recently used entry is dropped when a new slot shows up

    def _act_graphed(self, plan, key, launch):
        """A slot's actor step is the same launch sequence with the same pointers every iteration (the sampling
        stream advances on the device), so from its second use on it is replayed as a CUDA graph: the host cost
        of a step drops from ~10 launches + their TMA descriptor encodes to one replay."""
        graphs = plan.graphs
        g = graphs.pop(key, None)                             # re-inserted below: dict order = recency
        if not self.use_cuda_graph:
            return launch()
        if g is None:
            while len(graphs) >= self.MAX_ACT_GRAPHS:
                graphs.pop(next(iter(graphs)))
            graphs[key] = "seen"
            return launch()
        if g == "seen":
            dev = self._flat.device
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(dev)
            n0 = _native.launch_count()
            with torch.cuda.graph(graph):
                launch()
            n = _native.launch_count() - n0
            _native.lib().a2cb_add_launches(-n)               # capture recorded, did not run, them
            g = (graph, n)
        graphs[key] = g
        g[0].replay()
        _native.lib().a2cb_add_launches(g[1])

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
