"""ORACLE (test infrastructure only -- never imported by the product path).

Device-agnostic fp32 restatement of the VP-SDE tables and the predictor-corrector sampling loop as configured by the
reference (predictor 'ancestral_sampling', corrector 'none'):

  * VPSDE tables / marginal_prob   lib/diffusion/sde_lib.py:176-214
  * get_score_fn (std_scale=True)  lib/diffusion/models/utils.py:167-203
  * AncestralSamplingPredictor     lib/diffusion/sampling.py:212-236
  * pc_sampler unconditional loop  lib/diffusion/sampling.py:469-485
  * pc_sampler partial (cond) loop lib/diffusion/sampling.py:429-467

`model(x, labels)` is any callable; noise comes from `noise_fn(like)` so tests can share one stream between this
oracle, the reference and the CUDA path. Pinned against the reference by oracle/make_golden.py.
"""
import torch


class VPSDETables:
    def __init__(self, beta_min=0.1, beta_max=20.0, N=1000, device="cpu"):
        self.N, self.beta_0, self.beta_1, self.T = N, beta_min, beta_max, 1
        # sde_lib.py:189-195 -- linspace on the host, then the cumulative product on `device`
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N).to(device)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        mean = torch.exp(log_mean_coeff[:, None, None, None, None]) * x
        std = torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))
        return mean, std


def score_fn(sde, model, x, t):
    labels = t * (sde.N - 1)
    eps = model(x, labels)
    std = sde.sqrt_1m_alphas_cumprod[labels.long()]
    return -eps / std[:, None, None, None, None]


def ancestral_update(sde, model, x, t, noise_fn):
    timestep = (t * (sde.N - 1) / sde.T).long()
    beta = sde.discrete_betas[timestep]
    score = score_fn(sde, model, x, t)
    x_mean = (x + beta[:, None, None, None, None] * score) / torch.sqrt(1.0 - beta)[:, None, None, None, None]
    noise = noise_fn(x)
    x = x_mean + torch.sqrt(beta)[:, None, None, None, None] * noise
    return x, x_mean


def pc_sample_uncond(sde, model, x_init, grid_mask, noise_fn, eps=1e-3, n_iters=None, denoise=True):
    """sampling.py:469-485; n_iters truncates the loop (BASELINE config 1 uses the first 10 iterations)."""
    device = x_init.device
    timesteps = torch.linspace(sde.T, eps, sde.N, device=device)
    x = x_init * grid_mask
    x_mean = x
    total = sde.N - 1 if n_iters is None else n_iters
    for i in range(total):
        vec_t = torch.ones(x.shape[0], device=device) * timesteps[i]
        x, x_mean = x * grid_mask, x * grid_mask  # corrector 'none' returns (x, x), then the mask multiply
        x, x_mean = ancestral_update(sde, model, x, vec_t, noise_fn)
        x, x_mean = x * grid_mask, x_mean * grid_mask
    return x_mean if denoise else x


def pc_sample_partial(sde, model, x_init, grid_mask, partial, partial_mask, noise_fn, freeze_iters, eps=1e-3,
                      partial_channel=0, n_iters=None, denoise=True):
    """sampling.py:429-467 including the (B,B,...) broadcast of the initial re-noising (sampling.py:436-440)."""
    device = x_init.device
    c = partial_channel
    timesteps = torch.linspace(sde.T, eps, sde.N, device=device)
    x = x_init * grid_mask
    B = x.shape[0]
    vec_t = torch.ones(B, device=device) * timesteps[0]
    x[:, c] = partial[:, c] * grid_mask[:, c]
    pmean, pstd = sde.marginal_prob(x, vec_t)
    z = noise_fn(pmean[:, c])
    sampled = pmean[:, c] + pstd[:, None, None, None, None] * z  # broadcasts to (B,B,D,H,W) like the reference
    x[:, c] = (x[:, c] * (1 - partial_mask[:, c]) + sampled[:, c] * partial_mask[:, c]) * grid_mask[:, c]
    x_mean = x
    total = sde.N if n_iters is None else n_iters
    for i in range(total):
        vec_t = torch.ones(B, device=device) * timesteps[i]
        x, x_mean = x * grid_mask, x * grid_mask
        x, x_mean = ancestral_update(sde, model, x, vec_t, noise_fn)
        x, x_mean = x * grid_mask, x_mean * grid_mask
        if i != sde.N - 1 and i < freeze_iters:
            x[:, c] = (x[:, c] * (1 - partial_mask[:, c]) + partial[:, c] * partial_mask[:, c]) * grid_mask[:, c]
            x_mean[:, c] = (x_mean[:, c] * (1 - partial_mask[:, c]) + partial[:, c] * partial_mask[:, c]) * grid_mask[:, c]
            pmean, pstd = sde.marginal_prob(x, vec_t)
            z = noise_fn(pmean[:, c])
            sampled = pmean[:, c] + pstd[:, None, None, None] * z
            x[:, c] = (x[:, c] * (1 - partial_mask[:, c]) + sampled * partial_mask[:, c]) * grid_mask[:, c]
            x_mean[:, c] = x[:, c]
    return x_mean if denoise else x


def ddim_update(sde, model, x, t, tprev):
    """RSDE.discretize_ddim (lib/diffusion/sde_lib.py:113-140) with score_fn = raw noise prediction
    (get_score_fn(..., std_scale=False), models/utils.py:186-190): fp64 update, returns (x_new, x0_pred)."""
    step = (t * (sde.N - 1) / sde.T).long()
    step_prev = (tprev * (sde.N - 1) / sde.T).long()
    eps = model(x.float(), t.float() * (sde.N - 1))
    a1 = sde.sqrt_alphas_cumprod[step][:, None, None, None, None]
    a2 = sde.sqrt_1m_alphas_cumprod[step][:, None, None, None, None]
    a1p = sde.sqrt_alphas_cumprod[step_prev][:, None, None, None, None]
    a2p = sde.sqrt_1m_alphas_cumprod[step_prev][:, None, None, None, None]
    r1 = a1p.double() / a1.double()
    r2 = a2p.double() / a2.double()
    x0_scaled = x.double() - a2.double() * eps.double()
    noise_part = x - x0_scaled
    x0_pred = x0_scaled / a1
    x_new = r1.double() * x + (-r1 + r2.double()) * noise_part.double()
    return x_new, x0_pred
