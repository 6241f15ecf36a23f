"""Mesh post-ops behind the reference's names (nvdiffrec/lib/render/mesh.py:200-277, obj.py:165-216): smooth normals,
tangents and the OBJ writer, so that a generated `.npy` grid goes all the way to `mesh.obj` without the renderer stack.
The scatter-adds run on the GPU through the C ABI (`mdb_mesh_auto_normals`, `mdb_mesh_compute_tangents`)."""
import os

import numpy as np
import torch

from .. import _native


def auto_normals(v_pos, t_pos_idx):
    """v_pos fp32 [Nv,3], t_pos_idx int64 [F,3] (CUDA) -> (v_nrm [Nv,3], f_nrm [F,3]); t_nrm_idx == t_pos_idx."""
    L = _native.lib()
    v = v_pos.float().contiguous()
    f = t_pos_idx.long().contiguous()
    if not v.is_cuda:
        raise _native.NativeError("mesh post-ops run on the CUDA device only")
    Nv, F = v.shape[0], f.shape[0]
    v_nrm = torch.empty_like(v)
    f_nrm = torch.empty((F, 3), device=v.device, dtype=torch.float32)
    scratch = torch.empty(max(Nv, 1) * 3, device=v.device, dtype=torch.int64)
    _native.check(L.mdb_mesh_auto_normals(_native.ptr(v), _native.ptr(f), Nv, F, _native.ptr(v_nrm), _native.ptr(f_nrm),
                                          _native.ptr(scratch), _native.current_stream()))
    return v_nrm, f_nrm


def compute_tangents(v_pos, t_pos_idx, v_tex, t_tex_idx, v_nrm, t_nrm_idx):
    """-> v_tng fp32 [Nn,3]; t_tng_idx == t_nrm_idx."""
    L = _native.lib()
    v, uv, n = v_pos.float().contiguous(), v_tex.float().contiguous(), v_nrm.float().contiguous()
    tp, tt, tn = t_pos_idx.long().contiguous(), t_tex_idx.long().contiguous(), t_nrm_idx.long().contiguous()
    Nn, F = n.shape[0], tp.shape[0]
    out = torch.empty_like(n)
    scratch = torch.empty(max(Nn, 1) * 4, device=n.device, dtype=torch.int64)  # 3 int64 sums + 1 int32 count per vertex
    _native.check(L.mdb_mesh_compute_tangents(_native.ptr(v), _native.ptr(tp), _native.ptr(uv), _native.ptr(tt), _native.ptr(n),
                                              _native.ptr(tn), Nn, F, _native.ptr(out), _native.ptr(scratch), _native.current_stream()))
    return out


def write_obj(folder, v_pos, t_pos_idx, name="mesh.obj"):
    """Same text as obj.write_obj (positions and `i//` faces; the reference has texture/normal output commented out)."""
    v = v_pos.detach().cpu().numpy()
    f = t_pos_idx.detach().cpu().numpy()
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, name)
    with open(path, "w") as fh:
        fh.write("g default\n")
        fh.write("".join("v {} {} {} \n".format(p[0], p[1], p[2]) for p in v))
        fh.write("s 1 \ng pMesh1\nusemtl defaultMat\n")
        fi = f + 1
        fh.write("".join("f  {}// {}// {}//\n".format(a, b, c) for a, b, c in fi))
    return path
