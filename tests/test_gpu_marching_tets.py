"""Marching-tet parity on the real 64^3 tet grid: integer outputs bit-exact vs the golden vectors produced by the
REFERENCE DMTet class (oracle/make_golden.py) and vs the numpy oracle; vertices / uvs within rtol 1e-5."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import mt_oracle, synth

pytestmark = pytest.mark.gpu


def _grid():
    from meshdiffusion_b200.geometry import dmtet
    return dmtet.load_tet_grid(64)


@pytest.mark.parametrize("case,seed,noisy", [("sphere", 0, False), ("noisy", 1, True)])
def test_dmtet_call_matches_reference_golden(case, seed, noisy):
    from meshdiffusion_b200.geometry.dmtet import DMTet
    gold = load_golden("marching_tets_64.npz")
    verts, idx = _grid()
    sdf, pos = synth.synthetic_dmtet(verts, seed=seed, noisy=noisy)
    out = DMTet()(torch.tensor(pos).cuda(), torch.tensor(sdf).cuda(), torch.tensor(idx).long().cuda())
    v, f, uvs, uvi, f2t, vvi = [t.cpu().numpy() for t in out]
    assert f.dtype == np.int64 and uvi.dtype == np.int64 and f2t.dtype == np.int64 and vvi.dtype == np.int64
    assert np.array_equal(f, gold[case + "_faces"].astype(np.int64))
    assert np.array_equal(uvi, gold[case + "_uv_idx"].astype(np.int64))
    assert np.array_equal(f2t, gold[case + "_face_to_valid_tet"].astype(np.int64))
    assert np.array_equal(vvi, gold[case + "_valid_vert_idx"].astype(np.int64))
    assert np.allclose(v, gold[case + "_verts"], rtol=1e-5, atol=1e-6)
    assert tuple(uvs.shape) == tuple(gold[case + "_uvs_shape"])
    assert np.allclose(uvs[:64], gold[case + "_uvs_head"], rtol=1e-5, atol=1e-6)
    assert np.allclose(uvs[-64:], gold[case + "_uvs_tail"], rtol=1e-5, atol=1e-6)
    assert abs(uvs.astype(np.float64).sum() - gold[case + "_uvs_sum"][0]) < 1e-3 * abs(gold[case + "_uvs_sum"][0])


@pytest.mark.parametrize("case,seed,noisy", [("sphere", 0, False), ("noisy", 1, True)])
def test_dmtet_128_matches_reference_golden(case, seed, noisy):
    """BASELINE configs[3]/[4] grid size: the 128^3 tet grid (253 024 vertices, 1 387 746 tets). Integer outputs bit-exact
    against the REFERENCE DMTet class through length + sha256 (oracle/make_golden.py::golden_marching_tets_128), float
    outputs against sampled reference rows and sums."""
    import hashlib
    from meshdiffusion_b200.geometry import dmtet
    gold = load_golden("marching_tets_128.npz")
    verts, idx = dmtet.load_tet_grid(128)
    sdf, pos = synth.synthetic_dmtet(verts, seed=seed, noisy=noisy, res=128)
    out = dmtet.DMTet()(torch.tensor(pos).cuda(), torch.tensor(sdf).cuda(), torch.tensor(idx).long().cuda())
    names = ["verts", "faces", "uvs", "uv_idx", "face_to_valid_tet", "valid_vert_idx"]
    for n, t in zip(names, out):
        a = t.cpu().numpy()
        assert tuple(a.shape) == tuple(gold[f"{case}_{n}_shape"]), (n, a.shape)
        if a.dtype.kind in "iu":
            assert a.dtype == np.int64
            digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).digest(), dtype=np.uint8)
            assert np.array_equal(digest, gold[f"{case}_{n}_sha256"]), f"{n} is not bit-identical to the reference"
        else:
            assert np.allclose(a[gold[f"{case}_{n}_rows"]], gold[f"{case}_{n}_sample"], rtol=1e-5, atol=1e-6), n
            want = gold[f"{case}_{n}_sum"][0]
            assert abs(a.astype(np.float64).sum() - want) <= 1e-4 * max(abs(want), 1.0), n


def test_batched_extraction_matches_oracle():
    from meshdiffusion_b200.geometry.dmtet import MarchingTets
    verts, idx = _grid()
    B = 3
    sdfs, poss = zip(*[synth.synthetic_dmtet(verts, seed=10 + b, noisy=True) for b in range(B)])
    mt = MarchingTets(idx, verts.shape[0], max_batch=B)
    res = mt.extract(torch.tensor(np.stack(poss)).cuda(), torch.tensor(np.stack(sdfs)).cuda())
    for b in range(B):
        o = mt_oracle.marching_tets(poss[b], sdfs[b], idx)
        got = [t.cpu().numpy() for t in res[b]]
        for k in (1, 3, 4, 5):
            assert np.array_equal(got[k], o[k]), f"sample {b}: integer output {k} differs from the oracle"
        assert np.allclose(got[0], o[0], rtol=1e-5, atol=1e-6)


def test_degenerate_inputs():
    """All-inside and all-outside fields produce empty meshes (the reference returns empty tensors)."""
    from meshdiffusion_b200.geometry.dmtet import MarchingTets
    verts, idx = _grid()
    mt = MarchingTets(idx, verts.shape[0], max_batch=2)
    sdf = torch.stack([torch.ones(verts.shape[0]), -torch.ones(verts.shape[0])]).cuda()
    res = mt.extract(torch.tensor(verts).cuda(), sdf)
    for r in res:
        assert r[0].shape[0] == 0 and r[1].shape[0] == 0 and r[5].shape[0] == 0


def test_grid_to_mesh_pipeline():
    """eval.py:389-419 gather + dmtet.py:303 placement + marching tets, end to end against the oracle."""
    from meshdiffusion_b200.geometry import dmtet
    verts, idx = _grid()
    coords = dmtet.grid_coords_of_tet_vertices(verts)
    g = torch.Generator().manual_seed(0)
    grid = torch.randn(2, 4, 64, 64, 64, generator=g)
    sdf, pos = dmtet.grid_to_tet_inputs(grid.cuda(), coords.cuda(), torch.tensor(verts).cuda(), 64, mesh_scale=1.1, deform_scale=3.0)
    mt = dmtet.MarchingTets(idx, verts.shape[0], max_batch=2)
    res = mt.extract(pos, sdf)
    for b in range(2):
        s_o, p_o = mt_oracle.grid_to_tet_inputs(grid[b].numpy(), coords.numpy(), verts, 64, 1.1, 3.0)
        assert np.array_equal(sdf[b].cpu().numpy(), s_o)
        o = mt_oracle.marching_tets(p_o, s_o, idx)
        assert np.array_equal(res[b][1].cpu().numpy(), o[1])
        assert np.allclose(res[b][0].cpu().numpy(), o[0], rtol=1e-5, atol=1e-6)


def _dense_grad(gold, case, n_verts):
    gp = np.zeros((n_verts, 3), np.float32)
    gs = np.zeros(n_verts, np.float32)
    gp[gold[case + "_pos_rows"]] = gold[case + "_grad_pos"]
    gs[gold[case + "_sdf_rows"]] = gold[case + "_grad_sdf"]
    return gp, gs


@pytest.mark.parametrize("case,seed,noisy", [("sphere", 0, False), ("noisy", 1, True)])
def test_dmtet_gradients_match_reference_autograd(case, seed, noisy):
    """`verts` is differentiable like the reference's: d(sum(verts * W)) / d(pos_nx3, sdf_n) through `DMTet()(...)` equals
    torch autograd through the REFERENCE class (golden: oracle/make_golden.py::golden_marching_tets_grad), 1e-5 of the
    largest entry (fp32 sums in a different order)."""
    from meshdiffusion_b200.geometry.dmtet import DMTet
    gold = load_golden("marching_tets_64_grad.npz")
    verts, idx = _grid()
    sdf, pos = synth.synthetic_dmtet_grad_case(verts, seed=seed, noisy=noisy)
    p = torch.tensor(pos).cuda().requires_grad_(True)
    s = torch.tensor(sdf).cuda().requires_grad_(True)
    out = DMTet()(p, s, torch.tensor(idx).long().cuda())
    v = out[0]
    assert v.requires_grad and not out[1].requires_grad
    assert v.shape[0] == int(gold[case + "_n_verts"][0])
    W = torch.tensor(synth.mt_grad_weights(v.shape[0], seed)).cuda()
    (v * W).sum().backward()
    rp, rs = _dense_grad(gold, case, verts.shape[0])
    ep = np.abs(p.grad.cpu().numpy() - rp).max() / np.abs(rp).max()
    es = np.abs(s.grad.cpu().numpy() - rs).max() / np.abs(rs).max()
    print(f"marching-tet gradients {case}: pos {ep:.2e} sdf {es:.2e}")
    assert ep <= 1e-5 and es <= 1e-5


def test_batched_gradients_match_oracle_and_are_reproducible():
    """Batch of 3 with per-sample positions, then a shared vertex array (its gradient is the sum over the batch); a later
    extraction on the same handle must not disturb a pending backward; two runs are bitwise identical (gather, no atomics)."""
    from meshdiffusion_b200.geometry.dmtet import MarchingTets
    verts, idx = _grid()
    B = 3
    sdfs, poss = zip(*[synth.synthetic_dmtet_grad_case(verts, seed=20 + b, noisy=True) for b in range(B)])
    mt = MarchingTets(idx, verts.shape[0], max_batch=B)
    runs = []
    for _ in range(2):
        p = torch.tensor(np.stack(poss)).cuda().requires_grad_(True)
        s = torch.tensor(np.stack(sdfs)).cuda().requires_grad_(True)
        res = mt.extract(p, s)
        Ws = [synth.mt_grad_weights(res[b][0].shape[0], 50 + b) for b in range(B)]
        loss = sum((res[b][0] * torch.tensor(Ws[b]).cuda()).sum() for b in range(B))
        mt.extract(torch.tensor(poss[0]).cuda(), torch.tensor(np.stack(sdfs[:2])).cuda())  # reuses the handle's workspace
        loss.backward()
        runs.append((p.grad.clone(), s.grad.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    for b in range(B):
        gp, gs = mt_oracle.marching_tets_vertex_grad(poss[b], sdfs[b], idx, Ws[b])
        assert np.abs(runs[0][0][b].cpu().numpy() - gp).max() <= 1e-5 * np.abs(gp).max()
        assert np.abs(runs[0][1][b].cpu().numpy() - gs).max() <= 1e-5 * np.abs(gs).max()
    # shared vertex array
    p = torch.tensor(poss[0]).cuda().requires_grad_(True)
    s = torch.tensor(np.stack(sdfs)).cuda()
    res = mt.extract(p, s)
    Ws = [synth.mt_grad_weights(res[b][0].shape[0], 60 + b) for b in range(B)]
    sum((res[b][0] * torch.tensor(Ws[b]).cuda()).sum() for b in range(B)).backward()
    want = sum(mt_oracle.marching_tets_vertex_grad(poss[0], sdfs[b], idx, Ws[b])[0].astype(np.float64) for b in range(B))
    assert np.abs(p.grad.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()
