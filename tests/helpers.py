"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402


def tiny_config(name="res64", precision="tf32"):
    from configs import res64, res128
    cfg = (res128 if name == "res128" else res64).get_config()
    synth.apply_tiny(cfg, name)
    cfg.model.compute_dtype = precision
    return cfg


def full_config(name="res64", precision="tf32"):
    from configs import res64, res128
    cfg = (res128 if name == "res128" else res64).get_config()
    cfg.model.compute_dtype = precision
    return cfg


def build_model(cfg, device, state_seed):
    """Score network with the deterministic synthetic weights the golden vectors were generated with."""
    from meshdiffusion_b200.diffusion.models import utils as mutils
    cfg.device = torch.device(device)
    model = mutils.create_model(cfg)
    net = model.module
    sd = synth.synthetic_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, seed=state_seed)
    net.load_state_dict(sd)
    model.eval()  # inference engine; the training tests switch to .train() explicitly
    return model, sd


def rel_max(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def rel_l2(a, b):
    return ((a - b).double().pow(2).sum().sqrt() / b.double().pow(2).sum().sqrt()).item()


def load_golden(name):
    return np.load(os.path.join(GOLD, name))


def ddpm_loss(pred, noise, mask):
    """lib/diffusion/losses.py:69-78 (l2, masked)."""
    losses = torch.square(pred - noise) * mask
    losses = losses.reshape(losses.shape[0], -1).mean(dim=-1)
    return torch.mean(losses) / mask.sum() * mask.numel()


def grad_signature(name, g, seed=1234):
    """Same (norm, 4 random projections) signature as oracle/make_golden.py::grad_signature."""
    gen = torch.Generator().manual_seed(seed + sum(ord(c) for c in name))
    r = torch.randn(4, g.numel(), generator=gen, dtype=torch.float64)
    gd = g.detach().double().reshape(-1).cpu()
    return np.concatenate([[gd.norm().item()], (r @ gd).numpy()])
