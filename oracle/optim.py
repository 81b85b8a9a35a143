"""Oracle: the optimiser arithmetic of the update, written out explicitly.

The reference calls third-party torch (unpinned; pinned here to the installed
torch 2.11): optim.Adam (algo/ppo.py:32), optim.RMSprop (algo/a2c_acktr.py:30-31),
optim.SGD inside KFAC (algo/kfac.py:139-142) and nn.utils.clip_grad_norm_
(ppo.py:82-83, a2c_acktr.py:75-76).  These functions restate torch's single-tensor
update rules (torch/optim/{adam,rmsprop,sgd}.py, torch/nn/utils/clip_grad.py) so
that one optimiser step of the CUDA path can be checked without torch.optim.
"""
import math

import torch


def global_grad_norm(grads):
    return torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()


def clip_coef(total_norm, max_norm):
    """clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max=1); grads are ALWAYS multiplied."""
    return torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)


def adam_step(p, g, m, v, step, lr, eps, beta1=0.9, beta2=0.999):
    """In place; `step` is the 1-based count AFTER increment (torch _single_tensor_adam)."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def rmsprop_step(p, g, sq, lr, eps, alpha):
    sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
    p.addcdiv_(g, sq.sqrt().add_(eps), value=-lr)


def sgd_momentum_step(p, g, buf, first, lr, momentum):
    if first:
        buf.copy_(g)
    else:
        buf.mul_(momentum).add_(g)
    p.add_(buf, alpha=-lr)
