"""CPU: the oracle reproduces the golden vectors that oracle/make_golden.py produced from the REFERENCE modules."""
import numpy as np
import torch

from helpers import load_golden, tiny_config
from oracle import mt_oracle, sampler_oracle, synth, unet_oracle


def _tiny_sd(name, seed):
    from meshdiffusion_b200.diffusion.models import utils as mutils
    cfg = tiny_config(name)
    cfg.device = torch.device("cpu")
    net = mutils.create_model(cfg, use_parallel=False)
    sd = synth.synthetic_state_dict({k: v.detach() for k, v in net.state_dict().items()}, seed=seed)
    return cfg, sd


def test_unet_oracle_matches_reference_golden():
    for name in ("res64", "res128"):
        gold = load_golden(f"unet_tiny_{name}.npz")
        cfg, sd = _tiny_sd(name, int(gold["state_seed"]))
        assert np.abs(synth.state_checksum(sd) - gold["checksum"]).max() < 1e-6
        x, labels = synth.synthetic_inputs(cfg.data.image_size, 2, int(gold["input_seed"]), sd["mask"])
        with torch.no_grad():
            out = unet_oracle.unet_forward(sd, unet_oracle.arch_from_config(cfg), x, labels)
        assert np.abs(out.numpy() - gold["out"]).max() < 5e-5


def test_sampler_oracle_matches_reference_golden():
    gold = load_golden("sampler_tiny.npz")
    cfg, sd = _tiny_sd("res64", int(gold["state_seed"]))
    R, B, n_it = 16, 2, int(gold["n_iters"])
    sde = sampler_oracle.VPSDETables(cfg.model.beta_min, cfg.model.beta_max, cfg.model.num_scales)
    assert np.array_equal(sde.discrete_betas.numpy(), gold["betas"])
    assert np.array_equal(sde.sqrt_1m_alphas_cumprod.numpy(), gold["sqrt_1m_ac"])
    fn = lambda x, t: unet_oracle.unet_forward(sd, unet_oracle.arch_from_config(cfg), x, t)
    with torch.no_grad():
        torch.manual_seed(31)
        out = sampler_oracle.pc_sample_uncond(sde, fn, torch.randn(B, 4, R, R, R), sd["mask"].view(1, R, R, R), torch.randn_like, n_iters=n_it)
        assert np.abs(out.numpy() - gold["uncond"]).max() < 1e-4
        g = torch.Generator().manual_seed(41)
        partial = torch.sign(torch.randn(B, 4, R, R, R, generator=g))
        pmask = (torch.rand(1, 1, R, R, R, generator=g) < 0.5).float().expand(B, 4, R, R, R).contiguous()
        torch.manual_seed(32)
        out = sampler_oracle.pc_sample_partial(sde, fn, torch.randn(B, 4, R, R, R), sd["mask"].view(1, 1, R, R, R), partial, pmask,
                                               torch.randn_like, freeze_iters=3, n_iters=n_it)
        assert np.abs(out.numpy() - gold["partial"]).max() < 1e-4


def test_marching_tets_oracle_matches_reference_golden():
    from meshdiffusion_b200.geometry import dmtet
    gold = load_golden("marching_tets_64.npz")
    verts, idx = dmtet.load_tet_grid(64)
    for case, seed, noisy in (("sphere", 0, False), ("noisy", 1, True)):
        sdf, pos = synth.synthetic_dmtet(verts, seed=seed, noisy=noisy)
        v, f, uvs, uvi, f2t, vvi = mt_oracle.marching_tets(pos, sdf, idx)
        assert np.array_equal(f, gold[case + "_faces"])
        assert np.array_equal(uvi, gold[case + "_uv_idx"])
        assert np.array_equal(f2t, gold[case + "_face_to_valid_tet"])
        assert np.array_equal(vvi, gold[case + "_valid_vert_idx"])
        assert np.allclose(v, gold[case + "_verts"], rtol=1e-6, atol=1e-7)
        assert tuple(uvs.shape) == tuple(gold[case + "_uvs_shape"])


def _dense_grad(gold, case, n_verts):
    gp = np.zeros((n_verts, 3), np.float32)
    gs = np.zeros(n_verts, np.float32)
    gp[gold[case + "_pos_rows"]] = gold[case + "_grad_pos"]
    gs[gold[case + "_sdf_rows"]] = gold[case + "_grad_sdf"]
    return gp, gs


def test_marching_tets_gradient_oracle_matches_reference_autograd():
    """d(sum(verts * W)) / d(pos, sdf) of the numpy restatement vs torch autograd through the REFERENCE DMTet
    (oracle/make_golden.py::golden_marching_tets_grad)."""
    from meshdiffusion_b200.geometry import dmtet
    gold = load_golden("marching_tets_64_grad.npz")
    verts, idx = dmtet.load_tet_grid(64)
    for case, seed in (("sphere", 0), ("noisy", 1)):
        sdf, pos = synth.synthetic_dmtet_grad_case(verts, seed=seed, noisy=(case == "noisy"))
        W = synth.mt_grad_weights(int(gold[case + "_n_verts"][0]), seed)
        gp, gs = mt_oracle.marching_tets_vertex_grad(pos, sdf, idx, W)
        rp, rs = _dense_grad(gold, case, verts.shape[0])
        assert np.abs(gp - rp).max() <= 1e-5 * np.abs(rp).max()
        assert np.abs(gs - rs).max() <= 1e-5 * np.abs(rs).max()


def test_marching_tets_oracle_matches_reference_golden_128():
    """The R=128 grid (1.39 M tets): the oracle's integer outputs hash to the digests the REFERENCE DMTet class produced."""
    import hashlib
    from meshdiffusion_b200.geometry import dmtet
    gold = load_golden("marching_tets_128.npz")
    verts, idx = dmtet.load_tet_grid(128)
    sdf, pos = synth.synthetic_dmtet(verts, seed=0, noisy=False, res=128)
    out = mt_oracle.marching_tets(pos, sdf, idx)
    for n, a in zip(["verts", "faces", "uvs", "uv_idx", "face_to_valid_tet", "valid_vert_idx"], out):
        assert tuple(a.shape) == tuple(gold[f"sphere_{n}_shape"]), n
        if a.dtype.kind in "iu":
            digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).digest(), dtype=np.uint8)
            assert np.array_equal(digest, gold[f"sphere_{n}_sha256"]), n
        else:
            assert np.allclose(a[gold[f"sphere_{n}_rows"]], gold[f"sphere_{n}_sample"], rtol=1e-6, atol=1e-7), n


def test_grid_mask_from_tets_has_reference_population():
    from meshdiffusion_b200.geometry import dmtet
    m = dmtet.grid_mask_from_tets(64)
    assert int(m.sum()) == 30512 and m.shape == (64, 64, 64)  # 30 512 of 262 144 voxels
    assert int(dmtet.grid_mask_from_tets(128).sum()) == 253024


def test_ddim_oracle_matches_reference_golden():
    gold = load_golden("sampler_tiny.npz")
    cfg, sd = _tiny_sd("res64", int(gold["state_seed"]))
    R, B = 16, 2
    sde = sampler_oracle.VPSDETables(cfg.model.beta_min, cfg.model.beta_max, cfg.model.num_scales)
    fn = lambda x, t: unet_oracle.unet_forward(sd, unet_oracle.arch_from_config(cfg), x, t)
    g = torch.Generator().manual_seed(51)
    xd = torch.randn(B, 4, R, R, R, generator=g) * sd["mask"].view(1, R, R, R)
    with torch.no_grad():
        xn, x0 = sampler_oracle.ddim_update(sde, fn, xd, torch.full((B,), 0.64), torch.full((B,), 0.6084))
    assert np.abs(xn.numpy() - gold["ddim_x"]).max() < 2e-4 and np.abs(x0.numpy() - gold["ddim_x0"]).max() < 2e-4


def test_unet_oracle_backward_matches_reference_golden():
    """Autograd through the oracle reproduces the gradient signatures of the REFERENCE modules' loss.backward()."""
    from helpers import ddpm_loss, grad_signature
    gold = load_golden("unet_tiny_res64_grads.npz")
    cfg, sd = _tiny_sd("res64", int(gold["state_seed"]))
    R = cfg.data.image_size
    x, labels = synth.synthetic_inputs(R, 2, int(gold["input_seed"]), sd["mask"])
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
    osd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and k not in ("mask", "coords") else v) for k, v in sd.items()}
    loss = ddpm_loss(unet_oracle.unet_forward(osd, unet_oracle.arch_from_config(cfg), x, labels), noise, sd["mask"].view(1, 1, R, R, R))
    loss.backward()
    assert abs(loss.item() - float(gold["loss"])) < 1e-5 * float(gold["loss"])
    tot = float(gold["total_norm"])
    for n, sig in zip(gold["names"], gold["sig"]):
        got = grad_signature(str(n), osd[str(n)].grad)
        assert np.abs(got - sig).max() < 2e-4 * tot, n


def test_mesh_ops_oracle_matches_reference_golden():
    """auto_normals / compute_tangents restatement vs the reference functions' output on the synthetic sphere mesh."""
    from oracle import mesh_oracle
    from meshdiffusion_b200.geometry import dmtet
    gold = load_golden("mesh_ops_64.npz")
    verts, idx = dmtet.load_tet_grid(64)
    sdf, pos = synth.synthetic_dmtet(verts, seed=int(gold["seed"]), noisy=False)
    v, f, uvs, uv_idx, _, _ = mt_oracle.marching_tets(pos, sdf, idx)
    assert v.shape[0] == int(gold["n_verts"]) and f.shape[0] == int(gold["n_faces"])
    vn, fn = mesh_oracle.auto_normals(v, f)
    assert np.abs(vn - gold["v_nrm"]).max() < 1e-5
    assert np.abs(fn.astype(np.float64).sum(0) - gold["f_nrm_sum"]).max() < 1e-6
    vt = mesh_oracle.compute_tangents(v, f, uvs, uv_idx, vn, f)
    terr = np.abs(vt - gold["v_tng"]).max(1)
    assert np.percentile(terr, 99) < 1e-5 and (terr > 1e-3).mean() < 0.01
    text = mesh_oracle.obj_text(v[:3], f[:2])
    assert text.splitlines()[0] == "g default" and text.splitlines()[-1].startswith("f  ")
