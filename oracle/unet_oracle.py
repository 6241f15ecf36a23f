"""ORACLE (test infrastructure only -- never imported by the product path).

CPU/any-device fp32 restatement of the reference score network's forward pass, written functionally over a
state dict with the reference's own keys (without the DataParallel ``module.`` prefix).

Follows:
  * DDPMRes64.forward   lib/diffusion/models/ddpm_res64.py:126-199  (constructor order :57-123 for key numbering)
  * DDPMRes128.forward  lib/diffusion/models/ddpm_res128.py:137-215
  * ResnetBlockDDPM     lib/diffusion/models/layers.py:646-689
  * AttnBlock           layers.py:585-608      NIN layers.py:573-582
  * Upsample/Downsample layers.py:611-643
  * get_timestep_embedding layers.py:542-556

Pinned against the unmodified reference modules by oracle/make_golden.py (run in the authoring container where
/root/reference exists); the resulting vectors live in tests/golden/.  The reference ships no tests or golden
vectors of its own for this path, so parity is pinned by those generated vectors only.
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim, max_positions=10000):
    half = dim // 2
    scale = math.log(max_positions) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -scale)
    arg = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def _gn(sd, key, x):
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _nin(sd, key, x):
    # y[b,o,...] = sum_c x[b,c,...] W[c,o] + b[o]
    return torch.einsum("bcdhw,co->bodhw", x, sd[key + ".W"]) + sd[key + ".b"][None, :, None, None, None]


def _conv(sd, key, x, stride=1, padding=None):
    w = sd[key + ".weight"]
    if padding is None:
        padding = w.shape[-1] // 2
    return F.conv3d(x, w, sd[key + ".bias"], stride=stride, padding=padding)


def resblock(sd, pre, x, temb):
    cin = x.shape[1]
    h = _conv(sd, pre + "Conv_0", F.silu(_gn(sd, pre + "GroupNorm_0", x)))
    cout = h.shape[1]
    h = h + F.linear(F.silu(temb), sd[pre + "Dense_0.weight"], sd[pre + "Dense_0.bias"])[:, :, None, None, None]
    h = _conv(sd, pre + "Conv_1", F.silu(_gn(sd, pre + "GroupNorm_1", h)))  # eval mode: dropout is the identity
    if cin != cout:
        x = _nin(sd, pre + "NIN_0", x)
    return x + h


def attnblock(sd, pre, x):
    B, C, D, H, W = x.shape
    h = _gn(sd, pre + "GroupNorm_0", x)
    q = _nin(sd, pre + "NIN_0", h).reshape(B, C, -1)
    k = _nin(sd, pre + "NIN_1", h).reshape(B, C, -1)
    v = _nin(sd, pre + "NIN_2", h).reshape(B, C, -1)
    w = torch.einsum("bcq,bck->bqk", q, k) * (int(C) ** (-0.5))
    w = torch.softmax(w, dim=-1)
    o = torch.einsum("bqk,bck->bcq", w, v).reshape(B, C, D, H, W)
    return x + _nin(sd, pre + "NIN_3", o)


def downsample(sd, pre, x):
    return _conv(sd, pre + "Conv_0", F.pad(x, (0, 1, 0, 1, 0, 1)), stride=2, padding=0)


def upsample(sd, pre, x):
    B, C, D, H, W = x.shape
    return _conv(sd, pre + "Conv_0", F.interpolate(x, (D * 2, H * 2, W * 2), mode="nearest"))


def arch_from_config(config):
    """The structural hyper-parameters both reference networks read (ddpm_res64.py:46-53, ddpm_res128.py:48-55)."""
    name = config.model.name
    is128 = name.startswith("ddpm_res128")
    return dict(
        image_size=config.data.image_size, nf=config.model.nf, ch_mult=tuple(config.model.ch_mult),
        num_res_blocks=config.model.num_res_blocks, attn_resolutions=tuple(config.model.attn_resolutions),
        num_channels=config.data.num_channels, stem_ksize=5 if is128 else 3, use_pos_bias=not is128,
        level0_blocks=2 if is128 else config.model.num_res_blocks,
    )


def unet_forward(sd, arch, x, labels):
    """score_model(x, labels) in eval mode. sd: state dict (reference keys, no 'module.' prefix)."""
    nf, ch_mult, R = arch["nf"], arch["ch_mult"], arch["image_size"]
    nlev = len(ch_mult)
    blocks_at = lambda lvl: arch["level0_blocks"] if lvl == 0 else arch["num_res_blocks"]
    attn_at = lambda res: res in arch["attn_resolutions"]
    m = 0
    temb = timestep_embedding(labels, nf)
    temb = F.linear(temb, sd["all_modules.0.weight"], sd["all_modules.0.bias"])
    temb = F.linear(F.silu(temb), sd["all_modules.1.weight"], sd["all_modules.1.bias"])
    m = 2
    mask = sd["mask"]
    h = _conv(sd, "all_modules.2", x) + _conv(sd, "mask_layer", mask)
    if arch["use_pos_bias"]:
        # pos_layer(coords) with coords == 0 everywhere (ddpm_res64.py:74-78): a zero-padded conv of zeros = bias
        h = h + _conv(sd, "pos_layer", sd["coords"])
    m = 3
    hs = [h]
    for lvl in range(nlev):
        for _ in range(blocks_at(lvl)):
            h = resblock(sd, f"all_modules.{m}.", hs[-1], temb); m += 1
            if attn_at(h.shape[-1]):
                h = attnblock(sd, f"all_modules.{m}.", h); m += 1
            hs.append(h)
        if lvl != nlev - 1:
            hs.append(downsample(sd, f"all_modules.{m}.", hs[-1])); m += 1
    h = hs[-1]
    h = resblock(sd, f"all_modules.{m}.", h, temb); m += 1
    h = attnblock(sd, f"all_modules.{m}.", h); m += 1
    h = resblock(sd, f"all_modules.{m}.", h, temb); m += 1
    for lvl in reversed(range(nlev)):
        for _ in range(blocks_at(lvl) + 1):
            h = resblock(sd, f"all_modules.{m}.", torch.cat([h, hs.pop()], dim=1), temb); m += 1
        if attn_at(h.shape[-1]):
            h = attnblock(sd, f"all_modules.{m}.", h); m += 1
        if lvl != 0:
            h = upsample(sd, f"all_modules.{m}.", h); m += 1
    assert not hs
    h = F.silu(F.group_norm(h, 32, sd[f"all_modules.{m}.weight"], sd[f"all_modules.{m}.bias"], eps=1e-6)); m += 1
    h = _conv(sd, f"all_modules.{m}", h); m += 1
    return h


def nondegenerate_state_dict(sd, seed=0):
    """The reference initialises every Conv_1 / NIN_3 / head conv with scale 1e-10 (layers.py:90,593,662;
    ddpm_res64.py:121), so a random-init network outputs ~0 and relative error is meaningless. This redraws those
    tensors (and perturbs GroupNorm affines / biases) so activations are O(1) through the whole network.
    Deterministic in `seed`; operates on a copy."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    head = max(int(k.split(".")[1]) for k in sd if k.startswith("all_modules."))
    for k, v in sd.items():
        v = v.clone()
        if v.dtype.is_floating_point and k not in ("mask", "coords", "sigmas"):
            zero_init = k.endswith("Conv_1.weight") or k.endswith("NIN_3.W") or k == f"all_modules.{head}.weight"
            if zero_init:
                fan_in = v[0].numel() if v.dim() > 2 else v.shape[0]
                fan_out = v.shape[0] * (v[0, 0].numel() if v.dim() > 2 else 1) if v.dim() > 2 else v.shape[1]
                bound = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
                v = (torch.rand(v.shape, generator=g) * 2 - 1) * bound
            elif k.endswith(".bias") or k.endswith(".b"):
                if "GroupNorm" in k or k == f"all_modules.{head - 1}.bias":
                    v = torch.randn(v.shape, generator=g) * 0.1
                else:
                    v = torch.randn(v.shape, generator=g) * 0.02
            elif ("GroupNorm" in k and k.endswith(".weight")) or k == f"all_modules.{head - 1}.weight":
                v = torch.rand(v.shape, generator=g) + 0.5
        out[k] = v
    return out
