"""Oracle: GAIL discriminator update and reward pass (torch CPU + numpy).  TEST INFRASTRUCTURE ONLY.

Restates reference a2c_ppo_acktr/algo/gail.py:
  :11-29   Discriminator: trunk Linear(D, H) - Tanh - Linear(H, H) - Tanh - Linear(H, 1), default nn.Linear init,
           Adam with torch defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-8)
  :31-55   compute_grad_pen: alpha ~ U[0, 1) per row (CPU generator), mixup = alpha expert + (1 - alpha) policy,
           grad of the summed logits w.r.t. the mixup rows (double backward), lambda (||grad||_2 - 1)^2 mean
  :57-96   update: per (expert batch, policy batch) pair  BCE(expert, 1) + BCE(policy, 0) + penalty -> Adam step;
           returns the mean of the per-step losses
  :98-110  predict_reward: r = log s - log(1 - s), s = sigmoid(d); returns = returns * masks * gamma + r;
           RunningMeanStd update with the batch moments of `returns`; r / sqrt(var + 1e-8)
RunningMeanStd is stable_baselines3's (absent here, unpinned upstream): the published parallel-variance merge, restated
in `oracle/ref_shim.RunningMeanStd` -- the same class the reference runs on when the golden vectors are generated.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .ref_shim import RunningMeanStd


def init_params(input_dim, hidden_dim, seed=None):
    """Same RNG consumption as the reference constructor (three nn.Linear in order)."""
    if seed is not None:
        torch.manual_seed(seed)
    l0, l2, l4 = torch.nn.Linear(input_dim, hidden_dim), torch.nn.Linear(hidden_dim, hidden_dim), torch.nn.Linear(hidden_dim, 1)
    return {"trunk.0.weight": l0.weight.detach().clone(), "trunk.0.bias": l0.bias.detach().clone(),
            "trunk.2.weight": l2.weight.detach().clone(), "trunk.2.bias": l2.bias.detach().clone(),
            "trunk.4.weight": l4.weight.detach().clone(), "trunk.4.bias": l4.bias.detach().clone()}


def trunk(p, x):
    h = torch.tanh(F.linear(x, p["trunk.0.weight"], p["trunk.0.bias"]))
    h = torch.tanh(F.linear(h, p["trunk.2.weight"], p["trunk.2.bias"]))
    return F.linear(h, p["trunk.4.weight"], p["trunk.4.bias"])


def step_loss(p, expert, policy, alpha, lambda_=10.0):
    """(gail_loss + grad_pen) of one minibatch pair; `alpha` [B, 1] as drawn by the reference."""
    expert_d, policy_d = trunk(p, expert), trunk(p, policy)
    loss = F.binary_cross_entropy_with_logits(expert_d, torch.ones_like(expert_d)) + \
        F.binary_cross_entropy_with_logits(policy_d, torch.zeros_like(policy_d))
    a = alpha.expand_as(expert)
    mix = (a * expert + (1 - a) * policy).detach().requires_grad_(True)
    d = trunk(p, mix)
    grad = torch.autograd.grad(d, mix, torch.ones_like(d), create_graph=True, retain_graph=True, only_inputs=True)[0]
    return loss + lambda_ * (grad.norm(2, dim=1) - 1).pow(2).mean()


def policy_batches(obs_flat, act_flat, mini_batch_size):
    """feed_forward_generator(None, mini_batch_size=...) restricted to the two fields the discriminator reads
    (reference gail.py:60-67, storage.py:122-143): ONE randperm drawn when iteration starts, consecutive chunks,
    remainder dropped."""
    n = obs_flat.shape[0]
    perm = torch.randperm(n)
    for k in range(n // mini_batch_size):
        idx = perm[k * mini_batch_size:(k + 1) * mini_batch_size]
        yield obs_flat[idx], act_flat[idx]


def update(p, opt, expert_loader, obs_flat, act_flat, obsfilt):
    """p: dict of leaf tensors updated in place by `opt` (torch.optim.Adam over p.values()).  RNG consumption follows
    the reference: zip() starts the expert loader's iterator first, the permutation is drawn at the first policy
    batch, alpha (torch.rand(B, 1), global CPU generator) inside every step after the two BCE terms."""
    total, n = 0.0, 0
    for (es, ea), (ps, pa) in zip(expert_loader, policy_batches(obs_flat, act_flat, expert_loader.batch_size)):
        es = torch.FloatTensor(obsfilt(es.numpy(), update=False))
        expert, policy = torch.cat([es, ea], dim=1), torch.cat([ps, pa], dim=1)
        alpha = torch.rand(expert.size(0), 1)
        loss = step_loss(p, expert, policy, alpha)
        total += loss.item()
        n += 1
        opt.zero_grad()
        loss.backward()
        opt.step()
    return total / n


class RewardState(object):
    def __init__(self):
        self.returns = None
        self.ret_rms = RunningMeanStd(shape=())


def predict_reward(p, rs, state, action, gamma, masks, update_rms=True):
    with torch.no_grad():
        d = trunk(p, torch.cat([state, action], dim=1))
        s = torch.sigmoid(d)
        reward = s.log() - (1 - s).log()
        if rs.returns is None:
            rs.returns = reward.clone()
        if update_rms:
            rs.returns = rs.returns * masks * gamma + reward
            rs.ret_rms.update(rs.returns.cpu().numpy())
        return reward / np.sqrt(rs.ret_rms.var[0] + 1e-8)
