"""Generated grids -> meshes: the part of nvdiffrec/eval.py:386-447 that needs no renderer.

    python tools/npy_to_obj.py --sample_path <eval_dir>/0.npy --out_dir meshes [--resolution 64]
                               [--mesh_scale 1.1 --deform_scale 3.0]   (nvdiffrec/configs/res64.json:11,18)

Every grid of the .npy batch goes through the tet-vertex gather, marching tetrahedra, smooth normals and the OBJ writer
on the GPU (meshdiffusion_b200.geometry); one `<out_dir>/<index>/mesh.obj` per sample.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sample_path", required=True)
    ap.add_argument("--out_dir", required=True)
    ap.add_argument("--resolution", type=int, default=64)
    ap.add_argument("--mesh_scale", type=float, default=1.1)
    ap.add_argument("--deform_scale", type=float, default=3.0)
    args = ap.parse_args()
    from meshdiffusion_b200.geometry import dmtet, mesh_ops
    dev = torch.device("cuda:0")
    grids = torch.tensor(np.load(args.sample_path), dtype=torch.float32, device=dev)
    if grids.dim() == 4:
        grids = grids[None]
    R = args.resolution
    verts, idx = dmtet.load_tet_grid(R)
    v = torch.tensor(verts, device=dev)
    coords = dmtet.grid_coords_of_tet_vertices(v.cpu()).to(dev)
    sdf, pos = dmtet.grid_to_tet_inputs(grids, coords, v, R, args.mesh_scale, args.deform_scale)
    mt = dmtet.MarchingTets(idx, verts.shape[0], max_batch=grids.shape[0])
    for i, (mv, mf, uvs, uv_idx, _, _) in enumerate(mt.extract(pos, sdf)):
        v_nrm, _ = mesh_ops.auto_normals(mv, mf)
        path = mesh_ops.write_obj(os.path.join(args.out_dir, str(i)), mv, mf)
        print(f"{path}: {mv.shape[0]} vertices, {mf.shape[0]} faces")


if __name__ == "__main__":
    main()
