"""Oracle: loss terms of the policy update (differentiable torch expressions).

PPO:  reference algo/ppo.py:35-37 (advantage normalisation, UNBIASED std) and
      ppo.py:61-77 (clipped surrogate, clipped value loss).
A2C / ACKTR: reference algo/a2c_acktr.py:48-51 and the Fisher sampling loss :55-64.
"""
import torch


def normalized_advantages(returns, value_preds):
    adv = returns[:-1] - value_preds[:-1]
    return (adv - adv.mean()) / (adv.std() + 1e-5)


def ppo_losses(values, logp, old_values, returns, old_logp, adv, clip_param, use_clipped_value_loss=True):
    ratio = torch.exp(logp - old_logp)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param) * adv
    action_loss = -torch.min(surr1, surr2).mean()
    if use_clipped_value_loss:
        clipped = old_values + (values - old_values).clamp(-clip_param, clip_param)
        value_loss = 0.5 * torch.max((values - returns).pow(2), (clipped - returns).pow(2)).mean()
    else:
        value_loss = 0.5 * (returns - values).pow(2).mean()
    return value_loss, action_loss


def a2c_losses(values, logp, returns):
    adv = returns - values
    return adv.pow(2).mean(), -(adv.detach() * logp).mean()


def fisher_sample_loss(values, logp, value_noise):
    """-mean(logp) - mean((v - (v + noise).detach())^2)   (a2c_acktr.py:56-64)."""
    sample_values = values + value_noise
    return -logp.mean() - (values - sample_values.detach()).pow(2).mean()
