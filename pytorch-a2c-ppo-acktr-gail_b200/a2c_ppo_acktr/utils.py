"""Small utilities of the reference that sit on the update path.

`update_linear_schedule` (reference utils.py:46-50), `AddBias` (utils.py:32-43, the
log-std parameter holder of DiagGaussian) and `init` (utils.py:53-56).  The rendering /
log-directory helpers of the reference's utils.py are out of scope.
"""
import torch
import torch.nn as nn


class AddBias(nn.Module):
    """Holds a bias as a [n, 1] parameter named `_bias` (state_dict key `...logstd._bias`)."""

    def __init__(self, bias):
        super(AddBias, self).__init__()
        self._bias = nn.Parameter(bias.unsqueeze(1))

    def forward(self, x):
        b = self._bias.t().view(1, -1) if x.dim() == 2 else self._bias.t().view(1, -1, 1, 1)
        return x + b


def update_linear_schedule(optimizer, epoch, total_num_epochs, initial_lr):
    """Linear decay; the learning rate is read from param_groups at every step by the
    kernels' host wrapper, so writing it here takes effect immediately."""
    lr = initial_lr - (initial_lr * (epoch / float(total_num_epochs)))
    for group in optimizer.param_groups:
        group['lr'] = lr


def init(module, weight_init, bias_init, gain=1):
    weight_init(module.weight.data, gain=gain)
    bias_init(module.bias.data)
    return module
