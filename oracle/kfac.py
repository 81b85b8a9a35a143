"""Oracle: ACKTR's KFAC optimiser restated functionally on the CPU.

reference a2c_ppo_acktr/algo/kfac.py:16-242 and the ACKTR branch of algo/a2c_acktr.py:33-80.
The reference gathers layer inputs / output gradients with module hooks after rewriting the module
tree (SplitBias); here the same quantities are taken from an explicit functional forward pass
(`policy.base_forward` restated with captured intermediates) and `torch.autograd.grad`.
Closed forms:
  conv   aa = sum_rows p p^T / (B (oh ow)^2)   (p = unfold patches, (c,kh,kw) order);  gg = (B oh ow) sum_rows g g^T
  linear aa = x^T x / B;  gg = B g^T g;   bias aa = [[1]];  conv-bias gg = B s^T s, s = sum over pixels of g
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import losses


def cnn_forward_capture(p, obs):
    """CNNBase + Categorical head with every layer input kept and every pre-activation made a leaf
    of interest (for output gradients).  -> (value, logits, inputs dict, outputs dict)."""
    ins, outs = OrderedDict(), OrderedDict()
    x = obs.to(torch.float32) / 255.0
    for name, key, stride in (("conv1", "base.main.0", 4), ("conv2", "base.main.2", 2), ("conv3", "base.main.4", 1)):
        ins[name] = x
        y = F.conv2d(x, p[key + ".weight"], None, stride=stride)
        outs[name] = y                                       # output of the bias-free conv (== grad of AddBias output)
        x = F.relu(y + p[key + ".bias"].view(1, -1, 1, 1))
    x = x.reshape(x.shape[0], -1)
    ins["fc"] = x
    y = F.linear(x, p["base.main.7.weight"])
    outs["fc"] = y
    h = F.relu(y + p["base.main.7.bias"])
    ins["critic"] = ins["dist"] = h
    v = F.linear(h, p["base.critic_linear.weight"])
    outs["critic"] = v
    value = v + p["base.critic_linear.bias"]
    lg = F.linear(h, p["dist.linear.weight"])
    outs["dist"] = lg
    logits = lg + p["dist.linear.bias"]
    return value, logits, ins, outs


def mlp_forward_capture(p, obs):
    """MLPBase (model.py:198-229) + DiagGaussian head (distributions.py:76-95) with every layer input and every
    bias-free layer output kept.  -> (value, mean, logstd [B, A], inputs dict, outputs dict).  The logstd AddBias
    module (distributions.py:89-94) sees a zeros input and its own output's gradient: aa = [[1]], gg = B s^T s."""
    ins, outs = OrderedDict(), OrderedDict()
    x = obs.to(torch.float32)

    def lin(name, key, inp, act):
        ins[name] = inp
        y = F.linear(inp, p[key + ".weight"])
        outs[name] = y
        y = y + p[key + ".bias"]
        return torch.tanh(y) if act else y
    hc = lin("critic2", "base.critic.2", lin("critic0", "base.critic.0", x, True), True)
    ha = lin("actor2", "base.actor.2", lin("actor0", "base.actor.0", x, True), True)
    value = lin("critic", "base.critic_linear", hc, False)
    mean = lin("dist", "dist.fc_mean", ha, False)
    zeros = torch.zeros_like(mean)
    outs["logstd"] = zeros                                   # input of the AddBias module; its output gets the hook
    logstd = zeros + p["dist.logstd._bias"].t().view(1, -1)
    return value, mean, logstd, ins, outs


MLP_LAYERS = (("actor0", "base.actor.0"), ("actor2", "base.actor.2"), ("critic0", "base.critic.0"),
              ("critic2", "base.critic.2"), ("critic", "base.critic_linear"), ("dist", "dist.fc_mean"))

LAYERS = (("conv1", "base.main.0", (8, 4)), ("conv2", "base.main.2", (4, 2)), ("conv3", "base.main.4", (3, 1)),
          ("fc", "base.main.7", None), ("critic", "base.critic_linear", None), ("dist", "dist.linear", None))


def cov_a(a, conv):
    B = a.shape[0]
    if conv is not None:
        k, s = conv
        pt = F.unfold(a, kernel_size=k, stride=s)                       # [B, C*k*k, oh*ow]
        P = pt.shape[-1]
        rows = pt.transpose(1, 2).reshape(-1, pt.shape[1]) / P           # / oh / ow   (P = oh*ow; oh == ow here)
        return rows.t() @ (rows / B)
    return a.t() @ (a / B)


def cov_g(g, conv, bias):
    B = g.shape[0]
    if conv is not None and not bias:
        P = g.shape[2] * g.shape[3]
        rows = g.permute(0, 2, 3, 1).reshape(-1, g.shape[1]) * P
        g_ = rows * B
        return g_.t() @ (g_ / rows.shape[0])
    if conv is not None and bias:
        g = g.reshape(B, g.shape[1], -1).sum(-1)
    g_ = g * B
    return g_.t() @ (g_ / B)


class KFACState(object):
    def __init__(self, lr=0.25, momentum=0.9, stat_decay=0.99, kl_clip=0.001, damping=1e-2, Ts=1, Tf=10):
        self.lr, self.momentum, self.stat_decay, self.kl_clip, self.damping = lr, momentum, stat_decay, kl_clip, damping
        self.Ts, self.Tf, self.steps = Ts, Tf, 0
        self.m_aa, self.m_gg, self.Q_a, self.d_a, self.Q_g, self.d_g = {}, {}, {}, {}, {}, {}
        self.bufs = {}


def running(m, key, new, decay, first):
    if first:
        m[key] = new.clone()
    m[key] = m[key] * decay + new * (1 - decay) if not first else m[key]


def acktr_update(p, ro, hp, st):
    """p: dict of leaf tensors (requires_grad), updated in place.  st: KFACState."""
    if "dist.logstd._bias" in p:
        return acktr_update_mlp(p, ro, hp, st)
    T, N = ro["rewards"].shape[:2]
    B = T * N
    obs = ro["obs"][:-1].reshape(B, *ro["obs"].shape[2:])
    value, logits, ins, outs = cnn_forward_capture(p, obs)
    norm = logits - logits.logsumexp(-1, keepdim=True)
    actions = ro["actions"].reshape(B, 1)
    logp = norm.gather(-1, actions)
    ent = (-(norm.exp() * norm).sum(-1)).mean()
    values = value.view(T, N, 1)
    vl, al = losses.a2c_losses(values, logp.view(T, N, 1), ro["returns"][:-1])
    first = st.steps == 0
    if st.steps % st.Ts == 0:
        for name, key, conv in LAYERS:
            running(st.m_aa, key + ".weight", cov_a(ins[name].detach(), conv), st.stat_decay, first)
            running(st.m_aa, key + ".bias", torch.ones(1, 1), st.stat_decay, first)
        noise = torch.randn(values.size())
        fisher = losses.fisher_sample_loss(values, logp.view(T, N, 1), noise)
        gouts = torch.autograd.grad(fisher, [outs[n] for n, _, _ in LAYERS], retain_graph=True)
        for (name, key, conv), g in zip(LAYERS, gouts):
            running(st.m_gg, key + ".weight", cov_g(g, conv, False), st.stat_decay, first)
            running(st.m_gg, key + ".bias", cov_g(g, conv, True), st.stat_decay, first)
    total = vl * hp["value_loss_coef"] + al - ent * hp["entropy_coef"]
    keys = list(p.keys())
    grads = dict(zip(keys, torch.autograd.grad(total, [p[k] for k in keys])))
    return _kfac_step(p, grads, st, (vl, al, ent))


def acktr_update_mlp(p, ro, hp, st):
    """The same update for MLPBase + DiagGaussian (13 KFAC modules: 6 Linear weights, 6 biases, the logstd AddBias)."""
    from . import policy as o_pol
    T, N = ro["rewards"].shape[:2]
    B = T * N
    obs = ro["obs"][:-1].reshape(B, -1)
    value, mean, logstd, ins, outs = mlp_forward_capture(p, obs)
    actions = ro["actions"].reshape(B, -1)
    logp, ent = o_pol.gaussian_stats(mean, p["dist.logstd._bias"], actions)
    ent = ent.mean()
    values = value.view(T, N, 1)
    vl, al = losses.a2c_losses(values, logp.view(T, N, 1), ro["returns"][:-1])
    first = st.steps == 0
    if st.steps % st.Ts == 0:
        for name, key in MLP_LAYERS:
            running(st.m_aa, key + ".weight", cov_a(ins[name].detach(), None), st.stat_decay, first)
            running(st.m_aa, key + ".bias", torch.ones(1, 1), st.stat_decay, first)
        running(st.m_aa, "dist.logstd._bias", torch.ones(1, 1), st.stat_decay, first)
        noise = torch.randn(values.size())
        fisher = losses.fisher_sample_loss(values, logp.view(T, N, 1), noise)
        # the AddBias output is (zeros + bias): a per-row gradient w.r.t. the broadcast log-std, taken through a
        # per-row copy of the bias
        ls_rows = p["dist.logstd._bias"].t().view(1, -1).expand(B, -1).clone().requires_grad_(True)
        std = ls_rows.exp()
        lp_rows = (-((actions - mean) ** 2) / (2 * std * std) - ls_rows - math.log(math.sqrt(2 * math.pi))).sum(-1, keepdim=True)
        fisher_rows = losses.fisher_sample_loss(values, lp_rows.view(T, N, 1), noise)
        gouts = torch.autograd.grad(fisher, [outs[n] for n, _ in MLP_LAYERS], retain_graph=True)
        g_ls, = torch.autograd.grad(fisher_rows, [ls_rows], retain_graph=True)
        for (name, key), g in zip(MLP_LAYERS, gouts):
            running(st.m_gg, key + ".weight", cov_g(g, None, False), st.stat_decay, first)
            running(st.m_gg, key + ".bias", cov_g(g, None, False), st.stat_decay, first)
        running(st.m_gg, "dist.logstd._bias", cov_g(g_ls, None, False), st.stat_decay, first)
    total = vl * hp["value_loss_coef"] + al - ent * hp["entropy_coef"]
    keys = list(p.keys())
    grads = dict(zip(keys, torch.autograd.grad(total, [p[k] for k in keys])))
    return _kfac_step(p, grads, st, (vl, al, ent))


def _kfac_step(p, grads, st, loss_terms):
    vl, al, ent = loss_terms
    keys = list(p.keys())
    # ---- KFACOptimizer.step (kfac.py:190-242) ----
    updates = {}
    for k in keys:
        if st.steps % st.Tf == 0:
            st.d_a[k], st.Q_a[k] = torch.linalg.eigh(st.m_aa[k], UPLO="U")
            st.d_g[k], st.Q_g[k] = torch.linalg.eigh(st.m_gg[k], UPLO="U")
            st.d_a[k] = st.d_a[k] * (st.d_a[k] > 1e-6).float()
            st.d_g[k] = st.d_g[k] * (st.d_g[k] > 1e-6).float()
        g = grads[k]
        gm = g.reshape(g.shape[0], -1)
        v1 = st.Q_g[k].t() @ gm @ st.Q_a[k]
        v2 = v1 / (st.d_g[k].unsqueeze(1) * st.d_a[k].unsqueeze(0) + st.damping)
        updates[k] = (st.Q_g[k] @ v2 @ st.Q_a[k].t()).reshape(g.shape)
    vg = sum((updates[k] * grads[k] * st.lr * st.lr).sum() for k in keys)
    nu = min(1.0, math.sqrt(st.kl_clip / vg.item()))
    lr_in = st.lr * (1 - st.momentum)
    with torch.no_grad():
        for k in keys:
            d = updates[k] * nu
            if k not in st.bufs:
                st.bufs[k] = d.clone()
            else:
                st.bufs[k].mul_(st.momentum).add_(d)
            p[k].add_(st.bufs[k], alpha=-lr_in)
    st.steps += 1
    return vl.item(), al.item(), ent.item()
