"""Oracle: the actor-critic networks and action distributions, functionally.

Parameters are a dict keyed exactly like the reference `Policy.state_dict()`
(base.main.{0,2,4,7}.*, base.gru.*, base.actor.{0,2}.*, base.critic.{0,2}.*,
base.critic_linear.*, dist.linear.*, dist.fc_mean.*, dist.logstd._bias), so the
same tensors can be loaded into the reference, the oracle and the CUDA path.

Restates reference model.py:54-79 (act / get_value / evaluate_actions),
model.py:111-166 (GRU with episode-boundary masking), model.py:169-195 (CNNBase),
model.py:198-229 (MLPBase), distributions.py:18-44,59-95 and utils.py:32-43.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


class Spec(object):
    """obs_shape: (C,H,W) -> CNN trunk, (D,) -> MLP trunk.  action = ('discrete', n) | ('box', d)."""

    def __init__(self, obs_shape, action, recurrent=False, hidden_size=None):
        self.obs_shape = tuple(obs_shape)
        self.cnn = len(self.obs_shape) == 3
        if not self.cnn and len(self.obs_shape) != 1:
            raise NotImplementedError
        self.action_kind, self.num_actions = action
        self.recurrent = bool(recurrent)
        self.hidden = hidden_size if hidden_size is not None else (512 if self.cnn else 64)

    @property
    def hidden_state_size(self):
        return self.hidden if self.recurrent else 1


def _ortho(layer, gain):
    nn.init.orthogonal_(layer.weight.data, gain=gain)
    nn.init.constant_(layer.bias.data, 0)
    return layer


def init_params(spec, seed=None):
    """Draws parameters with the reference's initialisers in the reference's
    construction ORDER, so the global RNG stream is consumed identically
    (model.py:86-95 GRU first, then the trunk model.py:176-185 / 209-218, then the
    head distributions.py:63-69 / 80-84)."""
    if seed is not None:
        torch.manual_seed(seed)
    p = OrderedDict()

    def put(prefix, layer):
        p[prefix + ".weight"] = layer.weight.data.clone()
        p[prefix + ".bias"] = layer.bias.data.clone()

    gru = None
    if spec.recurrent:
        gru = nn.GRU(spec.hidden if spec.cnn else spec.obs_shape[0], spec.hidden)
        for name, w in gru.named_parameters():
            if "bias" in name:
                nn.init.constant_(w, 0)
            elif "weight" in name:
                nn.init.orthogonal_(w)
    g_relu = nn.init.calculate_gain("relu")
    if spec.cnn:
        put("base.main.0", _ortho(nn.Conv2d(spec.obs_shape[0], 32, 8, stride=4), g_relu))
        put("base.main.2", _ortho(nn.Conv2d(32, 64, 4, stride=2), g_relu))
        put("base.main.4", _ortho(nn.Conv2d(64, 32, 3, stride=1), g_relu))
        put("base.main.7", _ortho(nn.Linear(32 * 7 * 7, spec.hidden), g_relu))
        if gru is not None:
            for name, w in gru.named_parameters():
                p["base.gru." + name] = w.data.clone()
        put("base.critic_linear", _ortho(nn.Linear(spec.hidden, 1), 1))
    else:
        d_in = spec.hidden if spec.recurrent else spec.obs_shape[0]
        g = math.sqrt(2)
        if gru is not None:
            for name, w in gru.named_parameters():
                p["base.gru." + name] = w.data.clone()
        put("base.actor.0", _ortho(nn.Linear(d_in, spec.hidden), g))
        put("base.actor.2", _ortho(nn.Linear(spec.hidden, spec.hidden), g))
        put("base.critic.0", _ortho(nn.Linear(d_in, spec.hidden), g))
        put("base.critic.2", _ortho(nn.Linear(spec.hidden, spec.hidden), g))
        put("base.critic_linear", _ortho(nn.Linear(spec.hidden, 1), g))
    if spec.action_kind == "discrete":
        put("dist.linear", _ortho(nn.Linear(spec.hidden, spec.num_actions), 0.01))
    elif spec.action_kind == "box":
        put("dist.fc_mean", _ortho(nn.Linear(spec.hidden, spec.num_actions), 1))
        p["dist.logstd._bias"] = torch.zeros(spec.num_actions, 1)
    else:
        raise NotImplementedError
    return p


# ---------------------------------------------------------------------------
# trunks
# ---------------------------------------------------------------------------
def gru_cell(p, x, h):
    """One torch-GRU step, gate order (r, z, n)  [torch.nn.GRU semantics]."""
    gi = F.linear(x, p["base.gru.weight_ih_l0"], p["base.gru.bias_ih_l0"])
    gh = F.linear(h, p["base.gru.weight_hh_l0"], p["base.gru.bias_hh_l0"])
    H = h.shape[-1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def gru_forward(p, x, hxs, masks):
    """model.py:111-166.  The reference batches done-free segments through cuDNN/aten
    GRU; stepping every t with h * mask[t] is the same recurrence."""
    N = hxs.shape[0]
    if x.shape[0] == N:
        h = gru_cell(p, x, hxs * masks)
        return h, h
    T = x.shape[0] // N
    xs = x.view(T, N, -1)
    ms = masks.view(T, N, 1)
    # Like the reference (model.py:129-157): run the fused aten GRU over every stretch of steps in
    # which no environment resets, multiplying the carried state by the mask at the stretch start.
    w = [p["base.gru.weight_ih_l0"], p["base.gru.weight_hh_l0"], p["base.gru.bias_ih_l0"], p["base.gru.bias_hh_l0"]]
    cuts = [0] + [t for t in range(1, T) if bool((ms[t] == 0.0).any())] + [T]
    h = hxs.unsqueeze(0)
    outs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        y, h = torch._VF.gru(xs[a:b], h * ms[a].view(1, N, 1), w, True, 1, 0.0, False, False, False)
        outs.append(y)
    return torch.cat(outs, 0).view(T * N, -1), h.squeeze(0)


def base_forward(p, spec, obs, hxs, masks):
    """-> (value [B,1], actor_features [B,hidden], hxs)."""
    if spec.cnn:
        x = obs.to(torch.float32) / 255.0                                       # model.py:190
        x = F.relu(F.conv2d(x, p["base.main.0.weight"], p["base.main.0.bias"], stride=4))
        x = F.relu(F.conv2d(x, p["base.main.2.weight"], p["base.main.2.bias"], stride=2))
        x = F.relu(F.conv2d(x, p["base.main.4.weight"], p["base.main.4.bias"], stride=1))
        x = x.reshape(x.shape[0], -1)
        x = F.relu(F.linear(x, p["base.main.7.weight"], p["base.main.7.bias"]))
        if spec.recurrent:
            x, hxs = gru_forward(p, x, hxs, masks)
        return F.linear(x, p["base.critic_linear.weight"], p["base.critic_linear.bias"]), x, hxs
    x = obs
    if spec.recurrent:
        x, hxs = gru_forward(p, x, hxs, masks)
    hc = torch.tanh(F.linear(x, p["base.critic.0.weight"], p["base.critic.0.bias"]))
    hc = torch.tanh(F.linear(hc, p["base.critic.2.weight"], p["base.critic.2.bias"]))
    ha = torch.tanh(F.linear(x, p["base.actor.0.weight"], p["base.actor.0.bias"]))
    ha = torch.tanh(F.linear(ha, p["base.actor.2.weight"], p["base.actor.2.bias"]))
    return F.linear(hc, p["base.critic_linear.weight"], p["base.critic_linear.bias"]), ha, hxs


# ---------------------------------------------------------------------------
# heads
# ---------------------------------------------------------------------------
_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def categorical_stats(logits, action):
    """distributions.py:18-32 + torch.distributions.Categorical: -> (logp [B,1], entropy [B])."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    logp = norm.gather(-1, action.view(-1, 1).long())
    ent = -(norm.exp() * norm).sum(-1)
    return logp, ent


def gaussian_stats(mean, logstd_bias, action):
    """distributions.py:36-44,86-95 + torch.distributions.Normal: logstd is the AddBias
    parameter applied to zeros, std = exp(logstd)."""
    logstd = torch.zeros_like(mean) + logstd_bias.t().view(1, -1)
    std = logstd.exp()
    var = std ** 2
    logp = (-((action - mean) ** 2) / (2 * var) - std.log() - _LOG_SQRT_2PI).sum(-1, keepdim=True)
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + std.log()).sum(-1)
    return logp, ent


def head_forward(p, spec, feat):
    if spec.action_kind == "discrete":
        return F.linear(feat, p["dist.linear.weight"], p["dist.linear.bias"])
    return F.linear(feat, p["dist.fc_mean.weight"], p["dist.fc_mean.bias"])


def evaluate_actions(p, spec, obs, hxs, masks, action):
    """model.py:72-79 -> (value, action_log_probs, dist_entropy (mean), hxs)."""
    value, feat, hxs = base_forward(p, spec, obs, hxs, masks)
    out = head_forward(p, spec, feat)
    if spec.action_kind == "discrete":
        logp, ent = categorical_stats(out, action)
    else:
        logp, ent = gaussian_stats(out, p["dist.logstd._bias"], action)
    return value, logp, ent.mean(), hxs


def get_value(p, spec, obs, hxs, masks):
    return base_forward(p, spec, obs, hxs, masks)[0]


def act(p, spec, obs, hxs, masks, deterministic=False):
    """model.py:54-66.  Sampling draws from torch's global generator exactly like
    torch.distributions (multinomial on probs / normal(mean, std))."""
    value, feat, hxs = base_forward(p, spec, obs, hxs, masks)
    out = head_forward(p, spec, feat)
    if spec.action_kind == "discrete":
        norm = out - out.logsumexp(dim=-1, keepdim=True)
        probs = F.softmax(norm, dim=-1)
        if deterministic:
            action = probs.argmax(dim=-1, keepdim=True)
        else:
            action = torch.multinomial(probs, 1, True)
        logp, _ = categorical_stats(out, action)
    else:
        std = (torch.zeros_like(out) + p["dist.logstd._bias"].t().view(1, -1)).exp()
        action = out if deterministic else torch.normal(out, std)
        logp, _ = gaussian_stats(out, p["dist.logstd._bias"], action)
    return value, action, logp, hxs
