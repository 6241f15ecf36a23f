"""DDPM loss, optimiser and step function (reference: lib/diffusion/losses.py:26-141).

Same callables and semantics: Adam(lr, (beta1, 0.999), eps, weight_decay); linear warm-up on the MICRO-step counter
and `clip_grad_norm_`; loss = mean_b[mean((eps_theta - eps)^2 * mask)] * numel(mask) / sum(mask); `step_fn(state,
batch, clear_grad, update_param)` sums gradients over micro-batches (no division), advances `state['step']` and the
EMA every micro-step.
"""
import numpy as np
import torch
import torch.optim as optim

from .models import utils as mutils
from .sde_lib import VPSDE


def get_optimizer(config, params):
    if config.optim.optimizer != "Adam":
        raise NotImplementedError(f"Optimizer {config.optim.optimizer} not supported yet!")
    return optim.Adam(params, lr=config.optim.lr, betas=(config.optim.beta1, 0.999), eps=config.optim.eps,
                      weight_decay=config.optim.weight_decay)


def optimization_manager(config):
    def optimize_fn(optimizer, params, step, lr=config.optim.lr, warmup=config.optim.warmup,
                    grad_clip=config.optim.grad_clip):
        if warmup > 0:
            for group in optimizer.param_groups:
                group["lr"] = lr * np.minimum(step / warmup, 1.0)
        if grad_clip >= 0:
            torch.nn.utils.clip_grad_norm_(list(params), max_norm=grad_clip)
        optimizer.step()

    return optimize_fn


def get_ddpm_loss_fn(vpsde, train, mask=None, loss_type="l2"):
    if not isinstance(vpsde, VPSDE):
        raise TypeError("DDPM training only works for VPSDEs.")

    def loss_fn(model, batch):
        model_fn = mutils.get_model_fn(model, train=train)
        labels = torch.randint(0, vpsde.N, (batch.shape[0],), device=batch.device)
        a = vpsde.sqrt_alphas_cumprod.to(batch.device)[labels, None, None, None, None]
        s = vpsde.sqrt_1m_alphas_cumprod.to(batch.device)[labels, None, None, None, None]
        noise = torch.randn_like(batch)
        perturbed = (a * batch + s * noise) * mask
        pred = model_fn(perturbed, labels)
        if loss_type == "l2":
            losses = torch.square(pred - noise)
        elif loss_type == "l1":
            losses = torch.abs(pred - noise)
        else:
            raise NotImplementedError
        if mask is not None:
            losses = (losses * mask).reshape(losses.shape[0], -1).mean(dim=-1)
            return torch.mean(losses) / mask.sum() * np.prod(mask.size())
        return torch.mean(losses.reshape(losses.shape[0], -1).mean(dim=-1))

    return loss_fn


def get_step_fn(sde, train, optimize_fn=None, mask=None, loss_type="l2"):
    loss_fn = get_ddpm_loss_fn(sde, train, mask=mask, loss_type=loss_type)

    def step_fn(state, batch, clear_grad=True, update_param=True):
        model = state["model"]
        if train:
            optimizer = state["optimizer"]
            if clear_grad:
                optimizer.zero_grad()
            loss = loss_fn(model, batch)
            loss.backward()
            if update_param:
                optimize_fn(optimizer, model.parameters(), step=state["step"])
            state["step"] += 1
            state["ema"].update(model.parameters())
        else:
            with torch.no_grad():
                ema = state["ema"]
                ema.store(model.parameters())
                ema.copy_to(model.parameters())
                loss = loss_fn(model, batch)
                ema.restore(model.parameters())
        return {"loss": loss}

    return step_fn
