"""End-to-end through the reference's command line: `main_diffusion.py --mode=uncond_gen / cond_gen` write
float32 `[B,4,R,R,R]` .npy files that vanish outside the grid mask (BASELINE config 1 plumbing, config 5 conditioning)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _run(args, cwd):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main_diffusion.py")] + args, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return r


def test_uncond_gen_cli(tmp_path):
    out = os.path.join(tmp_path, "samples")
    _run([f"--config={ROOT}/configs/res64.py", "--mode=uncond_gen", f"--config.eval.eval_dir={out}",
          f"--config.eval.ckpt_path={tmp_path}/missing/checkpoint.pth", "--config.eval.batch_size=1",
          "--config.sampling.max_iters=10", "--config.model.compute_dtype=bf16"], cwd=str(tmp_path))
    x = np.load(os.path.join(out, "0.npy"))
    assert x.shape == (1, 4, 64, 64, 64) and x.dtype == np.float32 and np.isfinite(x).all()
    from meshdiffusion_b200.geometry.dmtet import grid_mask_from_tets
    m = grid_mask_from_tets(64).numpy()
    assert np.all(x[:, :, m == 0] == 0)
    assert abs((x != 0).mean() - 30512 / 262144) < 0.01  # 11.64 % of the voxels carry tet vertices


def test_cond_gen_cli_and_mesh_extraction(tmp_path):
    """Synthetic partial DMTet (sphere SDF, upper half visible) -> cond_gen -> grid->tet gather -> marching tets."""
    from meshdiffusion_b200.geometry import dmtet
    verts, idx = dmtet.load_tet_grid(64)
    v = torch.tensor(verts)
    partial = {"sdf": torch.sign(0.3 - v.norm(dim=1)), "vis": (v[:, 2] > 0).float()}
    ppath = os.path.join(tmp_path, "dmtet.pt")
    torch.save(partial, ppath)
    out = os.path.join(tmp_path, "cond")
    _run([f"--config={ROOT}/configs/res64.py", "--mode=cond_gen", f"--config.eval.eval_dir={out}",
          f"--config.eval.ckpt_path={tmp_path}/missing/checkpoint.pth", "--config.eval.batch_size=2",
          f"--config.eval.partial_dmtet_path={ppath}", f"--config.eval.tet_path={dmtet.tet_grid_path(64)}",
          "--config.sampling.max_iters=4", "--config.eval.freeze_iters=3", "--config.model.compute_dtype=bf16"], cwd=str(tmp_path))
    x = torch.from_numpy(np.load(os.path.join(out, "0.npy")))
    assert x.shape == (2, 4, 64, 64, 64) and torch.isfinite(x).all()
    coords = dmtet.grid_coords_of_tet_vertices(verts)
    sdf, pos = dmtet.grid_to_tet_inputs(x.cuda(), coords.cuda(), v.cuda(), 64, mesh_scale=1.1, deform_scale=3.0)
    meshes = dmtet.MarchingTets(idx, verts.shape[0], max_batch=2).extract(pos, sdf)
    for verts_b, faces_b, *_ in meshes:
        assert faces_b.shape[1] == 3 and faces_b.dtype == torch.int64
        if faces_b.numel():
            assert int(faces_b.max()) < verts_b.shape[0]


def test_train_cli_writes_reference_checkpoint_layout(tmp_path):
    """`--mode=train` (BASELINE config 3 plumbing): three optimiser steps on synthetic grids through the engine's
    forward + backward, then the checkpoint the reference's restore_checkpoint expects (utils.py:23-30)."""
    wd = os.path.join(tmp_path, "run")
    r = _run([f"--config={ROOT}/configs/res64.py", "--mode=train", f"--config.training.train_dir={wd}",
              "--config.data.synthetic=True", "--config.training.batch_size=1", "--config.training.n_iters=3",
              "--config.training.log_freq=1", "--config.training.snapshot_freq_for_preemption=2",
              "--config.training.snapshot_freq=100000"], cwd=str(tmp_path))
    ck = torch.load(os.path.join(wd, "checkpoints-meta", "checkpoint.pth"), map_location="cpu", weights_only=False)
    assert set(ck.keys()) == {"optimizer", "model", "ema", "step"}
    assert ck["step"] >= 2
    assert all(k.startswith("module.") for k in ck["model"])
    assert len(ck["ema"]["shadow_params"]) == 494
    losses = [float(l.split("training_loss:")[1]) for l in (r.stderr + r.stdout).splitlines() if "training_loss:" in l]
    assert len(losses) >= 3 and all(np.isfinite(losses))
    # loop bounds and names of the reference trainer (trainer.py:95,128-130): steps 0..n_iters, final checkpoint_<n_iters>.pth
    final = torch.load(os.path.join(wd, "checkpoints", "checkpoint_3.pth"), map_location="cpu", weights_only=False)
    assert final["step"] == 4 and set(final["optimizer"]["state"][2]) == {"step", "exp_avg", "exp_avg_sq"}
