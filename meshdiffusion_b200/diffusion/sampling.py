"""Predictor-corrector sampling for the VP-SDE score model.

Interface mirror of the reference's lib/diffusion/sampling.py: predictor / corrector registries (:33-80),
`get_sampling_fn` (:83-132), `get_pc_sampler` -> `pc_sampler(model, partial, partial_mask, partial_channel,
freeze_iters)` (:357-487). The configured pair (ancestral_sampling + none, configs/res64.py:22-23) runs on a fused
path: one native U-Net evaluation and ONE fused update kernel per step (score scaling, ancestral mean / noise and
both grid-mask multiplies), with the state resident in HBM and no host synchronisation inside the loop. Other
registered predictors / correctors use the same native network through `get_score_fn` with a few torch
elementwise ops around it.
"""
import abc
import ctypes

import numpy as np
import torch

from .. import _native
from . import sde_lib
from .models import utils as mutils

_PREDICTORS = {}
_CORRECTORS = {}


def register_predictor(cls=None, *, name=None):
    def _add(c):
        key = name if name is not None else c.__name__
        if key in _PREDICTORS:
            raise ValueError(f"Already registered model with name: {key}")
        _PREDICTORS[key] = c
        return c
    return _add if cls is None else _add(cls)


def register_corrector(cls=None, *, name=None):
    def _add(c):
        key = name if name is not None else c.__name__
        if key in _CORRECTORS:
            raise ValueError(f"Already registered model with name: {key}")
        _CORRECTORS[key] = c
        return c
    return _add if cls is None else _add(cls)


def get_predictor(name):
    return _PREDICTORS[name]


def get_corrector(name):
    return _CORRECTORS[name]


def get_sampling_fn(config, sde, shape, inverse_scaler, eps, grid_mask=None, return_traj=False):
    method = config.sampling.method.lower()
    if method == "pc":
        return get_pc_sampler(
            sde=sde, shape=shape,
            predictor=get_predictor(config.sampling.predictor.lower()),
            corrector=get_corrector(config.sampling.corrector.lower()),
            inverse_scaler=inverse_scaler, snr=config.sampling.snr, n_steps=config.sampling.n_steps_each,
            probability_flow=config.sampling.probability_flow, continuous=config.training.continuous,
            denoise=config.sampling.noise_removal, eps=eps, device=config.device, grid_mask=grid_mask,
            return_traj=return_traj,
            max_iters=config.sampling.get("max_iters", None), native_rng=config.sampling.get("native_rng", False),
            seed=config.get("seed", 42))
    if method == "ddim":
        return get_ddim_sampler(sde=sde, shape=shape, predictor=get_predictor("ddim"), inverse_scaler=inverse_scaler,
                                n_steps=config.sampling.n_steps_each, denoise=config.sampling.noise_removal, eps=eps,
                                device=config.device, grid_mask=grid_mask)
    raise ValueError(f"Sampler name {method} unknown.")


# ------------------------------------------------------------------------------------------------------------------
class Predictor(abc.ABC):
    def __init__(self, sde, score_fn, probability_flow=False):
        self.sde, self.score_fn, self.probability_flow = sde, score_fn, probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t):
        ...


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, n_steps):
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps

    @abc.abstractmethod
    def update_fn(self, x, t):
        ...


def _bcast(v):
    return v[:, None, None, None, None]


def _reverse_sde(sde, score_fn, x, t, probability_flow):
    """Drift / diffusion of the reverse-time SDE (sde_lib.py:100-107)."""
    drift, diffusion = sde.sde(x, t)
    score = score_fn(x, t)
    drift = drift - _bcast(diffusion) ** 2 * score * (0.5 if probability_flow else 1.0)
    return drift, (torch.zeros_like(diffusion) if probability_flow else diffusion)


def _reverse_discretize(sde, score_fn, x, t, probability_flow):
    """sde_lib.py:109-111."""
    f, G = sde.discretize(x, t)
    rev_f = f - _bcast(G) ** 2 * score_fn(x, t) * (0.5 if probability_flow else 1.0)
    return rev_f, (torch.zeros_like(G) if probability_flow else G)


@register_predictor(name="euler_maruyama")
class EulerMaruyamaPredictor(Predictor):
    def update_fn(self, x, t):
        dt = -1.0 / self.sde.N
        z = torch.randn_like(x)
        drift, diffusion = _reverse_sde(self.sde, self.score_fn, x, t, self.probability_flow)
        x_mean = x + drift * dt
        return x_mean + _bcast(diffusion) * np.sqrt(-dt) * z, x_mean


@register_predictor(name="reverse_diffusion")
class ReverseDiffusionPredictor(Predictor):
    def update_fn(self, x, t):
        f, G = _reverse_discretize(self.sde, self.score_fn, x, t, self.probability_flow)
        z = torch.randn_like(x)
        x_mean = x - f
        return x_mean + _bcast(G) * z, x_mean


@register_predictor(name="ancestral_sampling")
class AncestralSamplingPredictor(Predictor):
    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__(sde, score_fn, probability_flow)
        if not isinstance(sde, sde_lib.VPSDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        assert not probability_flow, "Probability flow not supported by ancestral sampling"

    def update_fn(self, x, t):
        sde = self.sde
        beta = sde.discrete_betas.to(t.device)[(t * (sde.N - 1) / sde.T).long()]
        score = self.score_fn(x, t)
        x_mean = (x + _bcast(beta) * score) / _bcast(torch.sqrt(1.0 - beta))
        return x_mean + _bcast(torch.sqrt(beta)) * torch.randn_like(x), x_mean


@register_predictor(name="none")
class NonePredictor(Predictor):
    def __init__(self, sde, score_fn, probability_flow=False):
        pass

    def update_fn(self, x, t):
        return x, x


@register_corrector(name="langevin")
class LangevinCorrector(Corrector):
    def update_fn(self, x, t):
        sde = self.sde
        alpha = sde.alphas.to(t.device)[(t * (sde.N - 1) / sde.T).long()]
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t)
            noise = torch.randn_like(x)
            grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
            noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
            step = (self.snr * noise_norm / grad_norm) ** 2 * 2 * alpha
            x_mean = x + _bcast(step) * grad
            x = x_mean + _bcast(torch.sqrt(step * 2)) * noise
        return x, x_mean


@register_corrector(name="ald")
class AnnealedLangevinDynamics(Corrector):
    def update_fn(self, x, t):
        sde = self.sde
        alpha = sde.alphas.to(t.device)[(t * (sde.N - 1) / sde.T).long()]
        std = sde.marginal_prob(x, t)[1]
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t)
            noise = torch.randn_like(x)
            step = (self.snr * std) ** 2 * 2 * alpha
            x_mean = x + _bcast(step) * grad
            x = x_mean + noise * _bcast(torch.sqrt(step * 2))
        return x, x_mean


@register_corrector(name="none")
class NoneCorrector(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        pass

    def update_fn(self, x, t):
        return x, x


# ------------------------------------------------------------------------------------------------------------------
def _native_net(model):
    from .models.ddpm import ScoreNet
    inner = getattr(model, "module", model)
    return inner if isinstance(inner, ScoreNet) else None


class _Cond:
    """Replacement conditioning of the partial branch (sampling.py:453-467 of the reference) in the form the update kernel
    takes it: channel `c` of `partial` / `partial_mask` (batch 1 = one grid shared by all samples, or one per sample) and
    the per-step marginal_prob scalars, computed with the reference's torch ops so they are bit-identical."""

    def __init__(self, sde, partial, partial_mask, c, timesteps, B):
        V = partial[0, 0].numel()
        self.c = int(c)
        self.partial = partial[:, c].to(torch.float32).contiguous()
        self.pmask = partial_mask[:, c].to(torch.float32).contiguous()
        for t in (self.partial, self.pmask):
            if t.shape[0] not in (1, B):
                raise ValueError("partial / partial_mask must have batch 1 or the sampling batch")
        self.pb = V if self.partial.shape[0] == B and B > 1 else 0
        self.mb = V if self.pmask.shape[0] == B and B > 1 else 0
        lmc = -0.25 * timesteps ** 2 * (sde.beta_1 - sde.beta_0) - 0.5 * timesteps * sde.beta_0  # sde_lib.py:211
        self.coefs = torch.exp(lmc).cpu().tolist()
        self.stds = torch.sqrt(1.0 - torch.exp(2.0 * lmc)).cpu().tolist()

    def struct(self, i=None, noise=None):
        s = _native.SamplerCondC()
        s.partial, s.partial_bstride = self.partial.data_ptr(), self.pb
        s.partial_mask, s.mask_bstride = self.pmask.data_ptr(), self.mb
        s.channel = self.c
        if i is not None:
            s.mean_coef, s.std = self.coefs[i], self.stds[i]
        s.noise = noise.data_ptr() if noise is not None else None
        return s


def _fused_update(eps, x, noise, mask_flat, beta, std, cond=None):
    """x, x_mean <- ancestral update (+ replacement conditioning when `cond` is given), in place on x; one kernel."""
    L = _native.lib()
    x_mean = torch.empty_like(x)
    B, C = x.shape[0], x.shape[1]
    V = x[0, 0].numel()
    _native.check(L.mdb_sampler_update(_native.ptr(eps), _native.ptr(x), _native.ptr(x_mean), _native.ptr(noise),
                                       _native.ptr(mask_flat), beta, std, V, C, B, 0, 0,
                                       ctypes.byref(cond) if cond is not None else None, _native.current_stream()))
    return x, x_mean


def _rank():
    import os
    return int(os.environ.get("RANK", "0"))


def get_pc_sampler(sde, shape, predictor, corrector, inverse_scaler, snr, n_steps=1, probability_flow=False,
                   continuous=False, denoise=True, eps=1e-3, device="cuda", grid_mask=None, return_traj=False,
                   max_iters=None, native_rng=False, seed=42):
    """Returns `pc_sampler(model, partial=None, partial_mask=None, partial_channel=0, freeze_iters=None)`.

    `max_iters` truncates the loop to its first iterations of the N-step schedule (used by the plumbing config;
    setting num_scales=10 instead would push beta above 1). `native_rng` draws the per-step noise inside the update
    kernel (Philox keyed by seed + rank, element index and step) and runs the whole loop inside the library -- the
    unconditional branch and the partial (`cond_gen`) branch alike.
    """
    fused = predictor is AncestralSamplingPredictor and corrector is NoneCorrector and not probability_flow

    def generic_step(model, x, vec_t):
        score_fn = mutils.get_score_fn(sde, model, train=False, continuous=continuous)
        c = NoneCorrector(sde, score_fn, snr, n_steps) if corrector is None else corrector(sde, score_fn, snr, n_steps)
        x, x_mean = c.update_fn(x, vec_t)
        x, x_mean = x * grid_mask, x_mean * grid_mask
        p = NonePredictor(sde, score_fn, probability_flow) if predictor is None else predictor(sde, score_fn, probability_flow)
        x, x_mean = p.update_fn(x, vec_t)
        return x * grid_mask, x_mean * grid_mask

    def compute_xzero(model, x, t, mask):
        step = (t * (sde.N - 1) / sde.T).long()
        a1, a2 = sde.sqrt_alphas_cumprod[step], sde.sqrt_1m_alphas_cumprod[step]
        eps_pred = model(x, t * torch.ones(shape[0], device=x.device))
        return ((x - a2 * eps_pred) / a1).clamp(-1, 1) * mask

    def pc_sampler(model, partial=None, partial_mask=None, partial_channel=0, freeze_iters=None):
        with torch.no_grad():
            if freeze_iters is None:
                freeze_iters = sde.N + 10
            c = partial_channel
            timesteps = torch.linspace(sde.T, eps, sde.N, device=device)
            x = sde.prior_sampling(shape).to(device)
            assert x.dim() == 5
            x = x * grid_mask
            B = shape[0]
            net = _native_net(model)
            binary_mask = bool(((grid_mask == 0) | (grid_mask == 1)).all())
            use_fused = fused and net is not None and binary_mask and x.is_cuda
            if use_fused:
                # per-step scalars, gathered once (no host sync inside the loop)
                idx = (timesteps * (sde.N - 1) / sde.T).long()
                betas = sde.discrete_betas.to(device)[idx].cpu().tolist()
                stds = sde.sqrt_1m_alphas_cumprod.to(device)[(timesteps * (sde.N - 1)).long()].cpu().tolist()
                labels_all = (timesteps * (sde.N - 1)).cpu().tolist()
                mask_flat = grid_mask.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
                assert mask_flat.numel() == x[0, 0].numel()
                x = x.contiguous()
            traj = []
            if use_fused:
                net._ensure_engine(B, x.device)
                net.sync_parameters()
                net._frozen = True  # nobody edits the weights inside the loop: skip per-step change detection
            try:
                return _run(model, net, x, use_fused, partial, partial_mask, c, freeze_iters, timesteps, B, traj,
                            betas if use_fused else None, stds if use_fused else None,
                            labels_all if use_fused else None, mask_flat if use_fused else None)
            finally:
                if use_fused:
                    net._frozen = False  # also on OOM / NativeError / KeyboardInterrupt inside the loop

    def _run(model, net, x, use_fused, partial, partial_mask, c, freeze_iters, timesteps, B, traj, betas, stds,
             labels_all, mask_flat):
        def step(x, i, cond=None):
            vec_t = torch.ones(B, device=device) * timesteps[i]
            if not use_fused:
                return generic_step(model, x, vec_t)
            eps_out = model(x, vec_t * (sde.N - 1))
            z = torch.randn_like(x)
            z2 = torch.randn_like(x[:, c]).contiguous() if cond is not None else None  # same draw order as the reference
            return _fused_update(eps_out, x, z, mask_flat, betas[i], stds[i], cond.struct(i, z2) if cond is not None else None)

        if partial is not None:
            assert partial.dim() == 5
            vec_t = torch.ones(B, device=device) * timesteps[0]
            x[:, c] = partial[:, c] * grid_mask[:, c]
            pmean, pstd = sde.marginal_prob(x, vec_t)
            # NB: (B,1,1,1,1) * (B,D,H,W) broadcasts to (B,B,D,H,W) in the reference (sampling.py:436-440)
            sampled = pmean[:, c] + pstd[:, None, None, None, None] * torch.randn_like(pmean[:, c])
            x[:, c] = (x[:, c] * (1 - partial_mask[:, c]) + sampled[:, c] * partial_mask[:, c]) * grid_mask[:, c]
            x_mean = x
            total = sde.N if max_iters is None else min(max_iters, sde.N)
            cond_until = min(freeze_iters, sde.N - 1)
            if use_fused:
                # the replacement + re-noising runs inside the update kernel (one launch per step); with native_rng the
                # whole loop runs inside the library
                cond = _Cond(sde, partial, partial_mask, c, timesteps, B)
                x = x.contiguous()
                if native_rng and not return_traj:
                    x_mean = _native_loop(net, x, mask_flat, labels_all, betas, stds, total, seed + _rank(), 0, cond, cond_until)
                else:
                    for i in range(total):
                        x, x_mean = step(x, i, cond if i < cond_until else None)
                total = 0
            for i in range(total):
                x, x_mean = step(x, i)
                if i != sde.N - 1 and i < freeze_iters:
                    keep, pm = 1 - partial_mask[:, c], partial_mask[:, c]
                    x[:, c] = (x[:, c] * keep + partial[:, c] * pm) * grid_mask[:, c]
                    x_mean[:, c] = (x_mean[:, c] * keep + partial[:, c] * pm) * grid_mask[:, c]
                    vec_t = torch.ones(B, device=device) * timesteps[i]
                    pmean, pstd = sde.marginal_prob(x, vec_t)
                    sampled = pmean[:, c] + pstd[:, None, None, None] * torch.randn_like(pmean[:, c])
                    x[:, c] = (x[:, c] * keep + sampled * pm) * grid_mask[:, c]
                    x_mean[:, c] = x[:, c]
        else:
            total = sde.N - 1 if max_iters is None else min(max_iters, sde.N - 1)
            if use_fused and native_rng and not return_traj:
                x_mean = _native_loop(net, x, mask_flat, labels_all, betas, stds, total, seed + _rank())
            else:
                x_mean = x
                for i in range(total):
                    x, x_mean = step(x, i)
                    if return_traj and i >= 700 and i % 10 == 0:
                        traj.append(compute_xzero(model, x, timesteps[i], grid_mask))
        if return_traj:
            return traj, sde.N * (n_steps + 1)
        return inverse_scaler(x_mean if denoise else x), sde.N * (n_steps + 1)

    return pc_sampler


def _native_loop(net, x, mask_flat, labels, betas, stds, total, seed, step0=0, cond=None, cond_until=0):
    """Steps step0 .. step0+total-1 of the loop inside the library (mdb_sampler_run): no Python between steps. The
    schedule lists are indexed by the GLOBAL step; `cond` (a _Cond) switches on the partial branch's replacement
    conditioning for the steps below `cond_until`."""
    L = _native.lib()
    B = x.shape[0]
    net._ensure_engine(B, x.device)
    net.sync_parameters()
    x_mean = torch.empty_like(x)
    eps_buf = torch.empty_like(x)
    labels_buf = torch.empty(B, device=x.device, dtype=torch.float32)
    arr = lambda v: (ctypes.c_float * total)(*v[step0:step0 + total])
    cs = cond.struct() if cond is not None else None
    _native.check(L.mdb_sampler_run(net._handle, _native.ptr(x), _native.ptr(x_mean), _native.ptr(mask_flat), arr(labels),
                                    arr(betas), arr(stds), total, B, int(seed), _native.ptr(eps_buf),
                                    _native.ptr(labels_buf), int(step0), ctypes.byref(cs) if cs is not None else None,
                                    arr(cond.coefs) if cond is not None else None, arr(cond.stds) if cond is not None else None,
                                    int(cond_until), _native.current_stream()))
    return x_mean


# ------------------------------------------------------------------------------------------------------------------
# DDIM (deterministic) sampling. The reference registers a 'ddim' predictor (sampling.py:249-257) on top of
# RSDE.discretize_ddim (sde_lib.py:113-140) and ships get_ddim_sampler (sampling.py:500-570), whose last line raises a
# NameError (`encode` is undefined). The update below is the same fp64 arithmetic; the sampler is the working version.
@register_predictor(name="ddim")
class DDIMPredictor(Predictor):
    def update_fn(self, x, t, tprev=None):
        sde = self.sde
        step = (t * (sde.N - 1) / sde.T).long()
        step_prev = (tprev * (sde.N - 1) / sde.T).long()
        eps = self.score_fn(x.float(), t.float())  # std_scale=False: the raw noise prediction
        a1 = _bcast(sde.sqrt_alphas_cumprod.to(x.device)[step])
        a2 = _bcast(sde.sqrt_1m_alphas_cumprod.to(x.device)[step])
        a1p = _bcast(sde.sqrt_alphas_cumprod.to(x.device)[step_prev])
        a2p = _bcast(sde.sqrt_1m_alphas_cumprod.to(x.device)[step_prev])
        r1 = a1p.double() / a1.double()
        r2 = a2p.double() / a2.double()
        x0_scaled = x.double() - a2.double() * eps.double()
        noise_part = x - x0_scaled
        x0_pred = x0_scaled / a1
        x_new = r1 * x + (-r1 + r2) * noise_part.double()
        return x_new, x0_pred


def get_ddim_sampler(sde, shape, predictor, inverse_scaler, n_steps=1, denoise=False, eps=1e-3, device="cuda",
                     grid_mask=None):
    def ddim_sampler(model, schedule="quad", num_steps=100, x0=None, partial=None, partial_mask=None, partial_channel=0):
        with torch.no_grad():
            x = (x0 if x0 is not None else sde.prior_sampling(shape).to(device)) * grid_mask
            c = partial_channel
            if partial is not None:
                x[:, c] = x[:, c] * (1 - partial_mask) + partial * partial_mask
            if schedule == "uniform":
                seq = list(range(0, sde.N, sde.N // num_steps))
            elif schedule == "quad":
                seq = [int(s) for s in np.linspace(0, np.sqrt(sde.N * 0.8), 100) ** 2]
            else:
                raise ValueError(f"unknown DDIM schedule {schedule}")
            timesteps = torch.tensor(seq) / sde.N
            score_fn = mutils.get_score_fn(sde, model, train=False, continuous=False, std_scale=False)
            pred = (predictor or DDIMPredictor)(sde, score_fn, False)
            x0_pred = x
            for i in reversed(range(1, len(timesteps))):
                vec_t = torch.ones(shape[0], device=device) * timesteps[i].to(device)
                vec_tprev = torch.ones(shape[0], device=device) * timesteps[i - 1].to(device)
                x, x0_pred = pred.update_fn(x, vec_t, vec_tprev)
                x, x0_pred = (x * grid_mask).float(), (x0_pred * grid_mask).float()
                if partial is not None:
                    x[:, c] = x[:, c] * (1 - partial_mask) + partial * partial_mask
                    x0_pred[:, c] = x0_pred[:, c] * (1 - partial_mask) + partial * partial_mask
            return inverse_scaler(x0_pred * grid_mask if denoise else x * grid_mask), sde.N * (n_steps + 1)

    return ddim_sampler
