"""Mesh post-ops on the GPU (mdb_mesh_auto_normals / mdb_mesh_compute_tangents through the Python mirror) vs the oracle
and the reference golden vectors; OBJ writer vs the oracle's literal restatement of obj.write_obj."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import mesh_oracle, mt_oracle, synth

pytestmark = pytest.mark.gpu


def _sphere_mesh(noisy):
    from meshdiffusion_b200.geometry import dmtet
    verts, idx = dmtet.load_tet_grid(64)
    sdf, pos = synth.synthetic_dmtet(verts, seed=3, noisy=noisy)
    got = dmtet.DMTet()(torch.tensor(pos).cuda(), torch.tensor(sdf).cuda(), torch.tensor(idx).long().cuda())
    return got  # verts, faces, uvs, uv_idx, ...


def test_auto_normals_and_tangents_match_reference_golden():
    from meshdiffusion_b200.geometry import mesh_ops
    gold = load_golden("mesh_ops_64.npz")
    v, f, uvs, uv_idx = _sphere_mesh(False)[:4]
    assert v.shape[0] == int(gold["n_verts"]) and f.shape[0] == int(gold["n_faces"])
    vn, fn = mesh_ops.auto_normals(v, f)
    e = np.abs(vn.cpu().numpy() - gold["v_nrm"]).max()
    print(f"normals vs reference: {e:.2e}")
    assert e < 1e-5
    assert np.abs(fn.double().sum(0).cpu().numpy() - gold["f_nrm_sum"]).max() < 1e-6
    vt = mesh_ops.compute_tangents(v, f, uvs, uv_idx, vn, f)
    terr = np.abs(vt.cpu().numpy() - gold["v_tng"]).max(1)
    print(f"tangents vs reference: p99 {np.percentile(terr, 99):.2e}, ill-conditioned {100 * (terr > 1e-3).mean():.2f} %")
    assert np.percentile(terr, 99) < 1e-5 and (terr > 1e-3).mean() < 0.01
    assert torch.isfinite(vt).all()


def test_normals_order_independent_and_degenerate_fallback():
    """Integer accumulation: permuting the faces does not change a single bit; unreferenced vertices get (0,0,1)."""
    from meshdiffusion_b200.geometry import mesh_ops
    v, f = _sphere_mesh(True)[:2]
    vn, _ = mesh_ops.auto_normals(v, f)
    perm = torch.randperm(f.shape[0], device=f.device, generator=torch.Generator(device="cuda").manual_seed(0))
    vn2, _ = mesh_ops.auto_normals(v, f[perm].contiguous())
    assert torch.equal(vn, vn2)
    v_extra = torch.cat([v, torch.zeros(2, 3, device=v.device)], 0)
    vn3, _ = mesh_ops.auto_normals(v_extra, f)
    assert torch.equal(vn3[-2:].cpu(), torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 1.0]]))
    ovn, _ = mesh_oracle.auto_normals(v.cpu().numpy(), f.cpu().numpy())
    err = np.abs(vn.cpu().numpy() - ovn).max(1)
    assert np.percentile(err, 99) < 1e-5


def test_write_obj_text(tmp_path):
    from meshdiffusion_b200.geometry import mesh_ops
    v, f = _sphere_mesh(False)[:2]
    path = mesh_ops.write_obj(str(tmp_path), v, f)
    assert open(path).read() == mesh_oracle.obj_text(v.cpu().numpy(), f.cpu().numpy())
