"""Marching tetrahedra on the sm_100a kernels, behind the reference's `DMTet()(pos_nx3, sdf_n, tet_fx4)` call
(nvdiffrec/lib/geometry/dmtet.py:32-163), plus the grid -> tet-vertex gather of nvdiffrec/eval.py:389-419 and the
vertex placement of dmtet.py:293-304. Integer outputs are int64 and bit-exact with the reference's ordering.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _native

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "tets")


def tet_grid_path(resolution, root=None):
    """`<root>/data/tets/<R>_tets_cropped.npz` like DMTetGeometry (dmtet.py:219), else the in-tree copy."""
    if root is not None:
        p = os.path.join(root, "data/tets/{}_tets_cropped.npz".format(resolution))
        if os.path.exists(p):
            return p
    return os.path.join(_DATA, "{}_tets_cropped.npz".format(resolution))


def load_tet_grid(resolution, root=None):
    t = np.load(tet_grid_path(resolution, root))
    return t["vertices"].astype(np.float32), t["indices"].astype(np.int32)


def grid_coords_of_tet_vertices(vertices):
    """Integer cubic-grid coordinate of every tet vertex (eval.py:391-397, evaler.py:187-195)."""
    v = torch.as_tensor(vertices)
    u = v.unique()
    dx = u[1] - u[0]
    return torch.round((v - v.min()) / dx).long()


def grid_mask_from_tets(resolution, root=None):
    """data/get_tet_mask.py: 1 where a tet vertex lands. Identical to the shipped data/grid_mask_<R>.pt."""
    verts, _ = load_tet_grid(resolution, root)
    c = grid_coords_of_tet_vertices(verts)
    mask = torch.zeros(resolution, resolution, resolution)
    mask[c[:, 0], c[:, 1], c[:, 2]] = 1.0
    return mask


class MarchingTets:
    """Static tet grid prepared once on the device; extracts meshes for batches of samples."""

    def __init__(self, tets, n_verts, max_batch=1):
        L = _native.lib()
        tets = np.ascontiguousarray(np.asarray(tets, dtype=np.int32))
        self.F, self.Nv, self.max_batch = tets.shape[0], int(n_verts), max_batch
        self._h = ctypes.c_void_p()
        _native.check(L.mdb_marching_tets_prepare(tets.ctypes.data_as(ctypes.c_void_p), self.F, self.Nv, max_batch,
                                                  ctypes.byref(self._h)))
        e, n = ctypes.c_int(), ctypes.c_int()
        L.mdb_marching_tets_info(self._h, ctypes.byref(e), ctypes.byref(n))
        self.n_edges, self.uv_n = e.value, n.value
        self._uvs = None

    def __del__(self):
        try:
            if self._h:
                _native.lib().mdb_marching_tets_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def uvs(self, device):
        if self._uvs is None:
            u = torch.empty(self.uv_n * self.uv_n * 4, 2, device=device, dtype=torch.float32)
            _native.check(_native.lib().mdb_marching_tets_uvs(self._h, _native.ptr(u), _native.current_stream()))
            self._uvs = u
        return self._uvs

    def extract(self, pos, sdf):
        """pos [B,Nv,3] or [Nv,3] (shared), sdf [B,Nv] fp32 cuda -> list of per-sample
        (verts, faces, uvs, uv_idx, face_to_valid_tet, valid_vert_idx)."""
        L = _native.lib()
        sdf = sdf.float().contiguous()
        B = sdf.shape[0]
        pos = pos.float().contiguous()
        stride = 0 if pos.dim() == 2 else self.Nv * 3
        counts = (ctypes.c_int * (3 * B))()
        stream = _native.current_stream()
        _native.check(L.mdb_marching_tets_count(self._h, _native.ptr(sdf), B, counts, stream))
        c = np.array(list(counts), dtype=np.int64).reshape(B, 3)
        off = np.concatenate([np.zeros((1, 3), np.int64), np.cumsum(c, 0)], 0)
        dev = sdf.device
        verts = torch.empty(int(off[-1, 0]), 3, device=dev, dtype=torch.float32)
        faces = torch.empty(int(off[-1, 1]), 3, device=dev, dtype=torch.int64)
        uv_idx = torch.empty_like(faces)
        f2t = torch.empty(int(off[-1, 1]), device=dev, dtype=torch.int64)
        vvi = torch.empty(int(off[-1, 2]), device=dev, dtype=torch.int64)
        # outputs are packed back to back in batch order: the library already holds those offsets on the device (NULL)
        _native.check(L.mdb_marching_tets_extract(self._h, _native.ptr(pos), stride, _native.ptr(sdf), B, _native.ptr(verts),
                                                  _native.ptr(faces), _native.ptr(uv_idx), _native.ptr(f2t), _native.ptr(vvi),
                                                  None, None, None, stream))
        uvs = self.uvs(dev)
        out = []
        for b in range(B):
            v0, v1 = off[b, 0], off[b + 1, 0]
            f0, f1 = off[b, 1], off[b + 1, 1]
            w0, w1 = off[b, 2], off[b + 1, 2]
            out.append((verts[v0:v1], faces[f0:f1], uvs, uv_idx[f0:f1], f2t[f0:f1], vvi[w0:w1]))
        return out


class DMTet:
    """Call-compatible with the reference class: `DMTet()(pos_nx3, sdf_n, tet_fx4)`."""

    def __init__(self):
        self._cache = {}

    def __call__(self, pos_nx3, sdf_n, tet_fx4):
        key = (tet_fx4.data_ptr(), tuple(tet_fx4.shape), pos_nx3.shape[0])
        if key not in self._cache:
            self._cache = {key: MarchingTets(tet_fx4.detach().cpu().numpy(), pos_nx3.shape[0], 1)}
        mt = self._cache[key]
        return mt.extract(pos_nx3, sdf_n[None])[0]


def grid_to_tet_inputs(grids, coords, vertices, grid_res, mesh_scale=1.0, deform_scale=1.0):
    """Batched eval.py:412-419 + dmtet.py:303: grids [B,4,R,R,R] -> (sdf [B,Nv], pos [B,Nv,3])."""
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    sdf = torch.sign(grids[:, 0, x, y, z])
    deform = grids[:, 1:, x, y, z].transpose(1, 2).clip(-1.0, 1.0)
    pos = vertices[None] * mesh_scale + 2 / (grid_res * 2) * deform * deform_scale
    return sdf.contiguous(), pos.contiguous()
