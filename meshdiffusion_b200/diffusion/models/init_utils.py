"""Synthetic "non-degenerate" initialisation for benchmarks and smoke tests.

The reference initialises every Conv_1, NIN_3 and the head convolution with scale 1e-10 (layers.py:90,593,662;
ddpm_res64.py:121), so a random-init network outputs ~0 and the activations of half the layers are numerically
trivial. For synthetic-weight benchmarks those tensors are redrawn at scale 1 and the GroupNorm affines perturbed,
which keeps every tensor O(1). Not used when a checkpoint is loaded.
"""
import torch

from .ddpm import variance_scaling_uniform


@torch.no_grad()
def random_init_nondegenerate(net, seed=1234):
    g = torch.Generator().manual_seed(seed)
    head = max(int(n.split(".")[1]) for n in net._names if n.startswith("all_modules."))
    for name in net._names:
        p = net._param(name)
        if name.endswith("Conv_1.weight") or name.endswith("NIN_3.W") or name == f"all_modules.{head}.weight":
            p.copy_(variance_scaling_uniform(tuple(p.shape), 1.0, generator=g).to(p.device))
        elif "GroupNorm" in name or name.startswith(f"all_modules.{head - 1}."):
            if name.endswith(".weight"):
                p.copy_((torch.rand(p.shape, generator=g) + 0.5).to(p.device))
            else:
                p.copy_((torch.randn(p.shape, generator=g) * 0.1).to(p.device))
        elif name.endswith(".bias") or name.endswith(".b"):
            p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
    return net
