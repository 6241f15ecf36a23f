"""`--mode=uncond_gen` / `--mode=cond_gen` drivers (reference: lib/diffusion/evaler.py:14-60, 134-211).

One process per GPU: under torchrun every rank loads the checkpoint, draws `eval.batch_size` samples with its own
seed and writes `<eval_dir>/<rank>.npy` (rank 0 writes `0.npy`, the reference's single-process file name); no
collective is involved. Output files are float32 `[B, 4, R, R, R]`, consumed by nvdiffrec/eval.py:400-417.
"""
import logging
import os

import numpy as np
import torch

from . import sampling, sde_lib
from .models import utils as mutils
from .models.ema import ExponentialMovingAverage
from .utils import restore_checkpoint


def _rank():
    return int(os.environ.get("RANK", "0"))


def load_grid_mask(resolution, device):
    """`./data/grid_mask_<R>.pt` (cwd-relative like evaler.py:38) if present, else derived from the tet grid."""
    path = "./data/grid_mask_{}.pt".format(resolution)
    if os.path.exists(path):
        return torch.load(path, map_location=device).to(device)
    from ..geometry.dmtet import grid_mask_from_tets
    return grid_mask_from_tets(resolution).to(device)


def _setup(config):
    device = config.device
    score_model = mutils.create_model(config)
    from . import losses
    optimizer = losses.get_optimizer(config, score_model.parameters())
    ema = ExponentialMovingAverage(score_model.parameters(), decay=config.model.ema_rate)
    state = dict(optimizer=optimizer, model=score_model, ema=ema, step=0)
    if config.training.sde.lower() != "vpsde":
        raise NotImplementedError(f"SDE {config.training.sde} unknown.")
    sde = sde_lib.VPSDE(beta_min=config.model.beta_min, beta_max=config.model.beta_max, N=config.model.num_scales,
                        device=device)
    return score_model, ema, state, sde


def uncond_gen(config, idx=None):
    idx = _rank() if idx is None else idx
    eval_dir = config.eval.eval_dir
    os.makedirs(eval_dir, exist_ok=True)
    torch.manual_seed(int(config.get("seed", 42)) + idx)
    score_model, ema, state, sde = _setup(config)
    R = config.data.image_size
    grid_mask = load_grid_mask(R, config.device).view(1, R, R, R)
    shape = (config.eval.batch_size, config.data.num_channels, R, R, R)
    sampling_fn = sampling.get_sampling_fn(config, sde, shape, lambda x: x, 1e-3, grid_mask=grid_mask)
    state = restore_checkpoint(config.eval.ckpt_path, state, device=config.device)
    ema.copy_to(score_model.parameters())
    logging.info("rank %d: sampling %d grids of %d^3", idx, shape[0], R)
    samples, _ = sampling_fn(score_model)
    out = os.path.join(eval_dir, f"{idx}.npy")
    np.save(out, samples.cpu().numpy())
    return out


def cond_gen(config, save_fname=None):
    save_fname = str(_rank()) if save_fname is None else save_fname
    eval_dir = config.eval.eval_dir
    os.makedirs(eval_dir, exist_ok=True)
    torch.manual_seed(int(config.get("seed", 42)) + _rank())
    score_model, ema, state, sde = _setup(config)
    device = config.device
    R = config.data.image_size
    grid_mask = load_grid_mask(R, device).view(1, 1, R, R, R)
    shape = (config.eval.batch_size, config.data.num_channels, R, R, R)
    sampling_fn = sampling.get_sampling_fn(config, sde, shape, lambda x: x, 1e-3, grid_mask=grid_mask)
    state = restore_checkpoint(config.eval.ckpt_path, state, device=device)
    ema.copy_to(score_model.parameters())

    partial = torch.load(config.eval.partial_dmtet_path, map_location=device)
    partial_sdf, partial_vis = partial["sdf"], partial["vis"]
    tet = np.load(config.eval.tet_path)
    from ..geometry.dmtet import grid_coords_of_tet_vertices
    c = grid_coords_of_tet_vertices(torch.tensor(tet["vertices"])).to(device)
    sdf_grid = torch.zeros(1, 1, R, R, R, device=device)
    sdf_grid[0, 0, c[:, 0], c[:, 1], c[:, 2]] = partial_sdf.to(device).float()
    vis_grid = torch.zeros(1, 1, R, R, R, device=device)
    vis_grid[0, 0, c[:, 0], c[:, 1], c[:, 2]] = partial_vis.to(device).float()
    samples, _ = sampling_fn(score_model, partial=sdf_grid, partial_mask=vis_grid,
                             freeze_iters=config.eval.freeze_iters)
    out = os.path.join(eval_dir, f"{save_fname}.npy")
    np.save(out, samples.cpu().numpy())
    return out
