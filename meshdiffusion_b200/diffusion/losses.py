"""Training-step host logic behind the reference's callables (lib/diffusion/losses.py:26-141): `get_optimizer`,
`optimization_manager`, `get_ddpm_loss_fn`, `get_step_fn`. Semantics kept:

* Adam(lr, (beta1, 0.999), eps, weight_decay); linear learning-rate warm-up on the MICRO-step counter, then
  `clip_grad_norm_`, then `optimizer.step()`.
* loss = mean_b[ mean_{c,v}((eps_theta - eps)^2 * mask) ] * numel(mask) / sum(mask) (l2; l1 uses |.|), with
  integer labels ~ U{0..N-1} and x_t = (sqrt(abar_t) x_0 + sqrt(1-abar_t) eps) * mask.
* `step_fn(state, batch, clear_grad, update_param)`: gradients of successive micro-batches add up (no division),
  `state['step']` and the EMA advance on EVERY micro-step, evaluation runs under the EMA weights.

The network forward and backward run inside the sm_100a engine; `loss.backward()` reaches it through one autograd node
(models/ddpm.py::_ScoreNetFn), so nothing here knows about the engine.
"""
import numpy as np
import torch

from .models import utils as mutils
from .sde_lib import VPSDE

_PENALTIES = {"l2": torch.square, "l1": torch.abs}


def get_optimizer(config, params):
    o = config.optim
    if o.optimizer != "Adam":
        raise NotImplementedError(f"Optimizer {o.optimizer} not supported yet!")
    return torch.optim.Adam(params, lr=o.lr, betas=(o.beta1, 0.999), eps=o.eps, weight_decay=o.weight_decay)


def optimization_manager(config):
    base_lr, base_warmup, base_clip = config.optim.lr, config.optim.warmup, config.optim.grad_clip

    def optimize_fn(optimizer, params, step, lr=base_lr, warmup=base_warmup, grad_clip=base_clip):
        if warmup > 0:
            scaled = lr * np.minimum(step / warmup, 1.0)
            for group in optimizer.param_groups:
                group["lr"] = scaled
        if grad_clip >= 0:
            torch.nn.utils.clip_grad_norm_(list(params), max_norm=grad_clip)
        optimizer.step()

    return optimize_fn


def get_ddpm_loss_fn(vpsde, train, mask=None, loss_type="l2"):
    if not isinstance(vpsde, VPSDE):
        raise TypeError("DDPM training only works for VPSDEs.")
    if loss_type not in _PENALTIES:
        raise NotImplementedError(loss_type)
    penalty = _PENALTIES[loss_type]

    def loss_fn(model, batch):
        dev = batch.device
        net = mutils.get_model_fn(model, train=train)
        t = torch.randint(0, vpsde.N, (batch.shape[0],), device=dev)
        signal = vpsde.sqrt_alphas_cumprod.to(dev)[t, None, None, None, None]
        sigma = vpsde.sqrt_1m_alphas_cumprod.to(dev)[t, None, None, None, None]
        eps = torch.randn_like(batch)
        x_t = (signal * batch + sigma * eps) * mask
        err = penalty(net(x_t, t) - eps)
        if mask is None:
            return err.reshape(err.shape[0], -1).mean(dim=-1).mean()
        per_sample = (err * mask).reshape(err.shape[0], -1).mean(dim=-1)
        return per_sample.mean() / mask.sum() * np.prod(mask.size())

    return loss_fn


def get_step_fn(sde, train, optimize_fn=None, mask=None, loss_type="l2"):
    loss_fn = get_ddpm_loss_fn(sde, train, mask=mask, loss_type=loss_type)

    def train_step(state, batch, clear_grad, update_param):
        model, optimizer = state["model"], state["optimizer"]
        if clear_grad:
            optimizer.zero_grad()
        loss = loss_fn(model, batch)
        loss.backward()
        if update_param:
            optimize_fn(optimizer, model.parameters(), step=state["step"])
        state["step"] += 1
        state["ema"].update(model.parameters())
        return loss

    def eval_step(state, batch):
        model, ema = state["model"], state["ema"]
        with torch.no_grad():
            ema.store(model.parameters())
            ema.copy_to(model.parameters())
            try:
                return loss_fn(model, batch)
            finally:
                ema.restore(model.parameters())

    def step_fn(state, batch, clear_grad=True, update_param=True):
        loss = train_step(state, batch, clear_grad, update_param) if train else eval_step(state, batch)
        return {"loss": loss}

    return step_fn
