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


def bind_host_to_device_node(device_index):
    """Host-side half of the rollout ingest (`RolloutStorage.stage_host`): pins the calling process to the CPU cores of
    the NUMA node the GPU hangs off, so that pinned rollout buffers allocated afterwards are local to the GPU's PCIe root
    (first-touch placement) -- with one rank per GPU every rank then pulls from its own memory controller instead of all
    of them from the node the launcher happened to start on.  Returns the CPU list it bound to, or None when the
    platform does not say (single-node hosts report numa_node = -1) or the container does not allow it."""
    import os
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            if int(f.read().strip()) < 0:
                return None
        with open(base + "/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
