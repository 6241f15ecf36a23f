"""BASELINE.json configs[4]: `cond_gen` at res64 with a synthetic partial DMTet (sdf = sign(0.3 - |v|), vis = v_z > 0), B = 4,
the full 1000-step partial branch with freeze_iters = 950, then the grid -> tet gather, marching tets and smooth normals.

    python tools/bench_cond_gen.py [--batch 4] [--precision bf16x3] [--steps 1000]

Everything goes through the public API: sampling.get_sampling_fn(config, ...)(model, partial, partial_mask, freeze_iters)
with `sampling.native_rng = True` (the loop then runs inside the library: mdb_sampler_run with the replacement
conditioning fused into the update kernel), geometry.dmtet.grid_to_tet_inputs / MarchingTets.extract, mesh_ops.auto_normals.
Prints one JSON line: wall seconds, sample-steps/s, samples/s, and the mesh stage.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "tf32", "bf16"])
    ap.add_argument("--steps", type=int, default=1000, help="iterations of the partial branch (1000 = the full run)")
    ap.add_argument("--freeze-iters", type=int, default=950)
    args = ap.parse_args()
    from configs import res64
    from meshdiffusion_b200.diffusion import sampling, sde_lib
    from meshdiffusion_b200.diffusion.models import utils as mutils
    from meshdiffusion_b200.diffusion.models.init_utils import random_init_nondegenerate
    from meshdiffusion_b200.geometry import dmtet, mesh_ops
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    cfg = res64.get_config()
    cfg.model.compute_dtype = args.precision
    cfg.device = dev
    cfg.sampling.native_rng = True
    if args.steps < 1000:
        cfg.sampling.max_iters = args.steps
    R, B = 64, args.batch
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    random_init_nondegenerate(model.module)
    model.eval()
    verts, idx = dmtet.load_tet_grid(R)
    v = torch.tensor(verts, device=dev)
    coords = dmtet.grid_coords_of_tet_vertices(v.cpu()).to(dev)
    mask = dmtet.grid_mask_from_tets(R).to(dev).view(1, 1, R, R, R)
    model.module.mask.data[:] = mask
    # evaler.py:181-201: scatter the per-vertex partial sdf / visibility into (1,1,R,R,R) grids
    sdf_grid = torch.zeros(1, 1, R, R, R, device=dev)
    sdf_grid[0, 0, coords[:, 0], coords[:, 1], coords[:, 2]] = torch.sign(0.3 - v.norm(dim=1))
    vis_grid = torch.zeros(1, 1, R, R, R, device=dev)
    vis_grid[0, 0, coords[:, 0], coords[:, 1], coords[:, 2]] = (v[:, 2] > 0).float()
    sde = sde_lib.VPSDE(cfg.model.beta_min, cfg.model.beta_max, cfg.model.num_scales, device=dev)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 4, R, R, R), lambda x: x, 1e-3, grid_mask=mask)
    cfg_w = cfg.sampling.get("max_iters", None)
    # warm-up: engine creation, weight packing, a few steps
    cfg.sampling.max_iters = 3
    sampling.get_sampling_fn(cfg, sde, (B, 4, R, R, R), lambda x: x, 1e-3, grid_mask=mask)(model, partial=sdf_grid, partial_mask=vis_grid, freeze_iters=args.freeze_iters)
    cfg.sampling.max_iters = cfg_w
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    samples, _ = fn(model, partial=sdf_grid, partial_mask=vis_grid, freeze_iters=args.freeze_iters)
    torch.cuda.synchronize()
    t_sample = time.perf_counter() - t0
    assert torch.isfinite(samples).all()
    n_steps = args.steps
    # the conditioned region of channel 0 must carry the partial sdf's sign pattern where it was frozen late
    t1 = time.perf_counter()
    sdf, pos = dmtet.grid_to_tet_inputs(samples, coords, v, R, mesh_scale=1.1, deform_scale=3.0)
    mt = dmtet.MarchingTets(idx, verts.shape[0], max_batch=B)
    meshes = mt.extract(pos, sdf)
    for mv, mf, *_ in meshes:
        if mf.shape[0]:
            mesh_ops.auto_normals(mv, mf)
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t1
    print(json.dumps({
        "workload": f"res64 cond_gen, batch {B}, {n_steps} iterations of the partial branch (freeze_iters {args.freeze_iters}), native loop + fused replacement conditioning, then marching tets + normals",
        "precision": args.precision, "sampling_seconds": t_sample, "sample_steps_per_s": B * n_steps / t_sample,
        "samples_per_s": B / t_sample if n_steps == 1000 else None, "ms_per_step": t_sample / n_steps * 1e3,
        "mesh_seconds": t_mesh, "tets_per_s": B * idx.shape[0] / t_mesh,
        "faces": [int(m[1].shape[0]) for m in meshes], "verts": [int(m[0].shape[0]) for m in meshes]}))


if __name__ == "__main__":
    main()
