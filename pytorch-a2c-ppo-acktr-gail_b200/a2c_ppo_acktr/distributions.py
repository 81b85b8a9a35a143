"""Action-distribution heads (parameter holders) of the drop-in Policy.

The reference builds torch.distributions objects per call (distributions.py:18-44,59-95);
here the head's matmul runs on the GEMM engine and log-prob / entropy / their gradients
come from the fused loss-epilogue kernel (csrc/loss.cu), so these modules only own the
parameters under the reference's names: `dist.linear.{weight,bias}` (Categorical, gain
0.01) and `dist.fc_mean.{weight,bias}` + `dist.logstd._bias` (DiagGaussian).
The reference's Bernoulli head is out of scope (not on the north-star
path and broken upstream at distributions.py:50).
"""
import torch
import torch.nn as nn

from .utils import AddBias, init


def _holder(layer):
    """Keeps only the parameters of a freshly initialised torch layer."""
    m = nn.Module()
    m.weight = nn.Parameter(layer.weight.data.clone())
    m.bias = nn.Parameter(layer.bias.data.clone())
    return m


class Categorical(nn.Module):
    def __init__(self, num_inputs, num_outputs):
        super(Categorical, self).__init__()
        self.linear = _holder(init(nn.Linear(num_inputs, num_outputs), nn.init.orthogonal_,
                                   lambda x: nn.init.constant_(x, 0), gain=0.01))
        self.num_outputs = num_outputs


class DiagGaussian(nn.Module):
    def __init__(self, num_inputs, num_outputs):
        super(DiagGaussian, self).__init__()
        self.fc_mean = _holder(init(nn.Linear(num_inputs, num_outputs), nn.init.orthogonal_,
                                    lambda x: nn.init.constant_(x, 0)))
        self.logstd = AddBias(torch.zeros(num_outputs))
        self.num_outputs = num_outputs
