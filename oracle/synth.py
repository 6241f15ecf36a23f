"""ORACLE support (test infrastructure only): deterministic synthetic weights / inputs shared by the golden-vector
generator (which also feeds them to the reference) and by the tests (which feed them to the CUDA path)."""
import math

import numpy as np
import torch


def apply_tiny(cfg, name):
    """A few-second version of the architecture that still exercises every layer type (SURVEY 8c 'cheap oracle shapes')."""
    cfg.data.image_size = 16
    cfg.model.nf = 32
    cfg.model.ch_mult = (1, 2)
    cfg.model.num_res_blocks = 2 if name == "res128" else 1
    cfg.model.attn_resolutions = (8,)
    return cfg


def synthetic_state_dict(template, seed):
    """Every tensor redrawn from one seeded CPU generator, in sorted-key order, with O(1) activations everywhere
    (the reference's own init zeroes Conv_1 / NIN_3 / head: layers.py:90,593,662; ddpm_res64.py:121)."""
    g = torch.Generator().manual_seed(seed)
    head = max(int(k.split(".")[1]) for k in template if k.startswith("all_modules."))
    out = {}
    for k in sorted(template.keys()):
        v = template[k]
        shape = tuple(v.shape)
        leaf = k.split(".")[-1]
        if k == "sigmas":
            out[k] = v.clone()
        elif k == "coords":
            out[k] = torch.zeros(shape)
        elif k == "mask":
            out[k] = (torch.rand(shape, generator=g) < 0.3).float()
        elif "GroupNorm" in k or k.startswith(f"all_modules.{head - 1}."):
            out[k] = torch.rand(shape, generator=g) + 0.5 if leaf == "weight" else torch.randn(shape, generator=g) * 0.1
        elif leaf in ("bias", "b"):
            out[k] = torch.randn(shape, generator=g) * 0.02
        else:
            if len(shape) > 2:
                rf = int(np.prod(shape[2:]))
                fan_in, fan_out = shape[1] * rf, shape[0] * rf
            else:
                fan_in, fan_out = shape[1], shape[0]
            bound = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
            out[k] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return {k: out[k] for k in template.keys()}


def state_checksum(sd):
    keys = sorted(k for k in sd if k != "sigmas")
    return np.array([float(sd[k].double().abs().sum()) for k in keys], np.float64)


def synthetic_inputs(R, batch, seed, mask):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 4, R, R, R, generator=g) * mask
    labels = torch.rand(batch, generator=g) * 999.0
    return x, labels


def synthetic_dmtet(vertices, seed, noisy, res=64):
    """sdf / vertex positions for marching-tet tests: a sphere of radius 0.3, optionally with
    sign noise (many disconnected components -> stresses the ordering rules) and random deformation."""
    rng = np.random.RandomState(seed)
    v = np.asarray(vertices, np.float32)
    r = np.linalg.norm(v, axis=1)
    sdf = (0.3 - r).astype(np.float32)
    deform = np.zeros_like(v)
    if noisy:
        flip = rng.rand(v.shape[0]) < 0.08
        sdf = np.where(flip, -sdf, sdf).astype(np.float32)
        sdf[rng.rand(v.shape[0]) < 0.01] = 0.0  # sign(0) = 0 is "outside" (occ = sdf > 0)
        deform = (rng.rand(*v.shape).astype(np.float32) - 0.5)
    sdf = np.sign(sdf).astype(np.float32) if noisy else sdf
    pos = (v * np.float32(1.1) + np.float32(2 / (res * 2)) * deform * np.float32(3.0)).astype(np.float32)
    return sdf, pos


def synthetic_dmtet_grad_case(vertices, seed, noisy):
    """Inputs of the marching-tet GRADIENT goldens: as synthetic_dmtet, but the noisy case keeps a continuous sdf (its signs
    with varying magnitudes; zeros stay zero) -- the sign() field of the sampling path has no gradient to check."""
    sdf, pos = synthetic_dmtet(vertices, seed=seed, noisy=noisy)
    if noisy:
        sdf = (sdf * (0.05 + np.random.RandomState(7).rand(sdf.shape[0]))).astype(np.float32)
    return sdf, pos


def mt_grad_weights(n_verts, seed):
    """dL/dverts of the marching-tet gradient goldens: L = sum(verts * W)."""
    return np.random.RandomState(100 + seed).randn(n_verts, 3).astype(np.float32)
