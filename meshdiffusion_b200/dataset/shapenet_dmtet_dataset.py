"""`ShapeNetDMTetDataset`: per-shape DMTet grids for `--mode=train`, with the item semantics of the reference loader
(lib/dataset/shapenet_dmtet_dataset.py:9-54). Items are bit-identical to the reference class on the same files and RNG
state (tests/test_host.py::test_dataset_items_match_reference_golden; golden produced by the reference class).

Pipeline of one item: load `[4, r, r, r]` grid (`.pt` or `.npy`) -> optional sign step -> optional jitter + grid-mask ->
right-pad every spatial axis to the mask resolution.

Two reference quirks are KEPT on purpose, each with an explicit opt-out:
  * sign step: the reference writes `datum[:, :1]`, which on a `[4, r, r, r]` tensor is the first X-slab of EVERY channel
    (channel 0 was meant). `fix_sign_axis=True` normalises channel 0 instead.
  * `.npy` files: the reference's branch raises NameError (numpy is never imported there); here the file is loaded.
"""
import json

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import Dataset

_JITTER = 0.01  # amplitude of the per-item translation noise on the three deformation channels


def _read_grid(path, extension):
    if extension == "npy":
        return torch.tensor(np.load(path))
    return torch.load(path, map_location="cpu")


def _shape_id(path, extension):
    """`..._<id>.<ext>` -> id, the key of the filter list."""
    stem = path.rstrip()[: -(len(extension) + 1)]
    return int(stem.split("_")[-1])


class ShapeNetDMTetDataset(Dataset):
    def __init__(self, root, grid_mask, deform_scale=1.0, aug=False, filter_meta_path=None, normalize_sdf=True,
                 extension="pt", fix_sign_axis=False):
        super().__init__()
        if extension not in ("pt", "npy"):
            raise AssertionError("extension must be 'pt' or 'npy'")
        with open(root, "r") as fh:
            paths = json.load(fh)
        if filter_meta_path is not None:
            with open(filter_meta_path, "r") as fh:
                self.filter_ids = json.load(fh)
            paths = [p for p in paths if _shape_id(p, extension) in self.filter_ids]
        self.fpath_list = paths
        self.extension = extension
        self.aug = aug
        self.normalize_sdf = normalize_sdf
        self.fix_sign_axis = fix_sign_axis
        self.deform_scale = deform_scale
        self.coeff = torch.tensor([1.0, 1.0, deform_scale, deform_scale, deform_scale]).view(-1, 1, 1, 1)  # kept: public attribute
        self.grid_mask = grid_mask.cpu()
        self.resolution = self.grid_mask.size(-1)

    def __len__(self):
        return len(self.fpath_list)

    def _sign_step(self, grid):
        region = grid[:1] if self.fix_sign_axis else grid[:, :1]
        s = torch.sign(region)
        s[s == 0] = 1.0
        region.copy_(s)

    def _augment(self, grid):
        r = grid.size(-1)
        occupied = grid[1:].abs().sum(dim=0, keepdim=True) != 0
        shift = (torch.rand(3)[:, None, None, None] - 0.5) * _JITTER  # one draw from the global CPU generator per item
        grid[1:] = grid[1:] + shift * occupied / (r / self.resolution)
        mask = self.grid_mask[0] if r >= self.resolution else self.grid_mask[0, :, :r, :r, :r]
        return grid * mask

    def __getitem__(self, idx):
        with torch.no_grad():
            grid = _read_grid(self.fpath_list[idx], self.extension)
            if self.normalize_sdf:
                self._sign_step(grid)
            if self.aug:
                grid = self._augment(grid)
        missing = self.resolution - grid.size(-1)
        if missing > 0:
            grid = F.pad(grid, (0, missing, 0, missing, 0, missing, 0, 0))
        return grid


def augment_on_device(raw, grid_mask, shifts, normalize_sdf=True, fix_sign_axis=False):
    """The per-item pipeline above for a whole batch on the GPU: raw [B,4,r,r,r] grids as loaded (device
    tensor), grid_mask [1,1,R,R,R], shifts [B,3] ~ U(0,1) (the `torch.rand(3)` draw of each item). Returns [B,4,R,R,R].
    Same arithmetic, same order as `ShapeNetDMTetDataset.__getitem__` with aug=True, so given the items' draws the result
    is bit-identical to stacking the loader's items (tests/test_host.py::test_on_device_augmentation_matches_items)."""
    grid = raw.clone()
    R, r = grid_mask.size(-1), grid.size(-1)
    if normalize_sdf:
        region = grid[:, :1] if fix_sign_axis else grid[:, :, :1]   # the reference's quirk: first X-slab of every channel
        s = torch.sign(region)
        s[s == 0] = 1.0
        region.copy_(s)
    occupied = grid[:, 1:].abs().sum(dim=1, keepdim=True) != 0
    shift = (shifts.to(grid.device)[:, :, None, None, None] - 0.5) * _JITTER
    grid[:, 1:] = grid[:, 1:] + shift * occupied / (r / R)
    mask = grid_mask.to(grid.device)[0] if r >= R else grid_mask.to(grid.device)[0, :, :r, :r, :r]
    grid = grid * mask
    if R - r > 0:
        grid = F.pad(grid, (0, R - r, 0, R - r, 0, R - r, 0, 0, 0, 0))
    return grid

