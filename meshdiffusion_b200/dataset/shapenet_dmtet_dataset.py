"""`ShapeNetDMTetDataset` with the reference's constructor and item semantics
(lib/dataset/shapenet_dmtet_dataset.py:9-54): a JSON list of per-shape grid files ([4, r, r, r] tensors, `.pt` or `.npy`),
optional id filter, sign "normalisation", jitter augmentation of the deformation channels, grid-mask multiply and
right-padding to the mask resolution.

Two reference quirks, kept on purpose unless the caller opts out:
  * the sign normalisation writes `datum[:, :1]` on a [4, r, r, r] tensor, i.e. it replaces the FIRST X-SLAB OF EVERY
    CHANNEL by its sign (the intent was channel 0). `fix_sign_axis=True` normalises channel 0 instead.
  * the `.npy` branch of the reference raises NameError (numpy is never imported there); here it loads the file.
Augmentation draws `torch.rand(3)` from the global CPU generator exactly like the reference, so a seeded run produces the
same items.
"""
import json

import numpy as np
import torch
from torch.utils.data import Dataset


class ShapeNetDMTetDataset(Dataset):
    def __init__(self, root, grid_mask, deform_scale=1.0, aug=False, filter_meta_path=None, normalize_sdf=True,
                 extension="pt", fix_sign_axis=False):
        super().__init__()
        self.fpath_list = json.load(open(root, "r"))
        self.deform_scale = deform_scale
        self.normalize_sdf = normalize_sdf
        self.fix_sign_axis = fix_sign_axis
        self.coeff = torch.tensor([1.0, 1.0, deform_scale, deform_scale, deform_scale]).view(-1, 1, 1, 1)
        self.aug = aug
        self.grid_mask = grid_mask.cpu()
        self.resolution = self.grid_mask.size(-1)
        self.extension = extension
        assert self.extension in ["pt", "npy"]
        if filter_meta_path is not None:
            self.filter_ids = json.load(open(filter_meta_path, "r"))
            ids = [int(x.rstrip().split("_")[-1][:-len(self.extension) - 1]) for x in self.fpath_list]
            self.fpath_list = [p for p, i in zip(self.fpath_list, ids) if i in self.filter_ids]

    def __len__(self):
        return len(self.fpath_list)

    def __getitem__(self, idx):
        with torch.no_grad():
            if self.extension == "pt":
                datum = torch.load(self.fpath_list[idx], map_location="cpu")
            else:
                datum = torch.tensor(np.load(self.fpath_list[idx]))
            if self.normalize_sdf:
                target = datum[:1] if self.fix_sign_axis else datum[:, :1]
                sign = torch.sign(target)
                sign[sign == 0] = 1.0
                target.copy_(sign)
            if self.aug:
                nonempty = (datum[1:].abs().sum(dim=0, keepdim=True) != 0)
                datum[1:] = datum[1:] + (torch.rand(3)[:, None, None, None] - 0.5) * 0.01 * nonempty / (datum.size(-1) / self.resolution)
                r = datum.size(-1)
                datum = datum * (self.grid_mask[0, :, :r, :r, :r] if r < self.resolution else self.grid_mask[0])
        if datum.size(-1) < self.resolution:
            d = self.resolution - datum.size(-1)
            datum = torch.nn.functional.pad(datum, (0, d, 0, d, 0, d, 0, 0))
        return datum
