"""Optimiser-side training kernels vs the reference's torch formulation (losses.py:26-85, ema.py:43-64) on the GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ddpm_loss_and_gradient():
    from meshdiffusion_b200 import train_ops
    g = torch.Generator(device="cuda").manual_seed(0)
    B, C, R = 3, 4, 16
    pred = torch.randn(B, C, R, R, R, device="cuda", generator=g, requires_grad=True)
    noise = torch.randn(B, C, R, R, R, device="cuda", generator=g)
    mask = (torch.rand(1, 1, R, R, R, device="cuda", generator=g) < 0.12).float()
    losses = torch.square(pred - noise) * mask
    ref = torch.mean(losses.reshape(B, -1).mean(dim=-1)) / mask.sum() * np.prod(mask.size())
    ref.backward()
    loss, grad = train_ops.ddpm_loss(pred.detach(), noise, mask, want_grad=True)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert torch.allclose(grad, pred.grad, rtol=1e-5, atol=1e-8)


class _RefEMA:
    """ExponentialMovingAverage.update exactly as the reference writes it (ema.py:43-64), in torch ops."""

    def __init__(self, params, decay):
        self.decay, self.num_updates = decay, 0
        self.shadow_params = [p.detach().clone() for p in params if p.requires_grad]

    def update(self, params):
        self.num_updates += 1
        d = min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            for s, p in zip(self.shadow_params, [p for p in params if p.requires_grad]):
                s.sub_((1.0 - d) * (s - p))


def test_fused_adam_ema_matches_torch():
    """train_ops.FusedAdam (clip coefficient + Adam + EMA in one native pass) vs clip_grad_norm_ + torch.optim.Adam + the
    reference EMA, over tensors that exercise the vector path, the scalar tail and an unaligned chunk; the optimiser
    state_dict must be interchangeable with torch.optim.Adam's."""
    from meshdiffusion_b200 import train_ops
    from meshdiffusion_b200.diffusion.models.ema import ExponentialMovingAverage
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(128, 64, 3, 3, 3), (512,), (256, 128), (7,), (8192 * 3 + 5,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    opt = torch.optim.Adam(ref_p, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    ema_ref = _RefEMA(ref_p, 0.9999)
    ema_ours = ExponentialMovingAverage(our_p, decay=0.9999)
    fused = train_ops.FusedAdam(our_p, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    for it in range(4):
        grads = [torch.randn(s, device="cuda", generator=g) * (3.0 if it % 2 else 0.01) for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad = gr.clone(); q.grad = gr.clone()
        lr = 2e-5 * min((it + 1) / 5000, 1.0)
        for group in opt.param_groups + fused.param_groups:
            group["lr"] = lr
        tn_ref = torch.nn.utils.clip_grad_norm_(ref_p, max_norm=1.0)
        opt.step()
        ema_ref.update(ref_p)
        tn = fused.grad_norm_coef(1.0)
        assert fused.step(ema=ema_ours) is True
        if it == 2:  # an accumulation micro-step afterwards: EMA alone (native mdb_ema_update), no Adam
            ema_ours.update(our_p)
            ema_ref.update(ref_p)
        assert abs(tn.item() - tn_ref.item()) <= 1e-5 * tn_ref.item()
        for p, q in zip(ref_p, our_p):
            assert torch.allclose(q, p, rtol=2e-6, atol=1e-9)
        for s, e in zip(ema_ref.shadow_params, ema_ours.shadow_params):
            assert torch.allclose(e, s, rtol=2e-6, atol=1e-9)
    assert ema_ours.num_updates == ema_ref.num_updates
    # same state layout as torch.optim.Adam: a stock optimiser loads ours and continues identically
    sd = fused.state_dict()
    ref_sd = opt.state_dict()
    assert set(sd) == set(ref_sd) and set(sd["state"][0]) == set(ref_sd["state"][0])
    assert float(sd["state"][0]["step"]) == float(ref_sd["state"][0]["step"]) == 4.0
    stock = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in our_p], lr=2e-5)
    stock.load_state_dict(sd)
    assert torch.allclose(stock.state_dict()["state"][0]["exp_avg"], ref_sd["state"][0]["exp_avg"], rtol=1e-5, atol=1e-9)


def test_product_step_fn_matches_torch_adam_and_reference_ema():
    """The PRODUCT training step (losses.get_step_fn as trainer.py wires it: native perturb + loss node, engine backward,
    FusedAdam with the EMA folded in) vs the same step with the stock optimiser pieces (torch.optim.Adam + clip_grad_norm_ +
    the reference EMA): 3 steps, parameters and EMA equal to 2e-6 after EVERY step. The stock run's parameters are
    re-aligned to the product run's after each comparison: a 1-ulp fp32 difference flips bf16 roundings of packed weights,
    which would otherwise feed back through the next gradient (that is the engine's operand precision, not the optimiser)."""
    from helpers import build_model, tiny_config
    from meshdiffusion_b200 import train_ops
    from meshdiffusion_b200.diffusion import losses, sde_lib
    from meshdiffusion_b200.diffusion.models.ema import ExponentialMovingAverage
    cfg = tiny_config("res64", "bf16")
    cfg.model.dropout = 0.0
    cfg.optim.warmup = 2
    cfg.optim.lr = 1e-3
    R, B = 16, 2
    sde = sde_lib.VPSDE(cfg.model.beta_min, cfg.model.beta_max, cfg.model.num_scales, device="cuda:0")
    runs = []
    for fused in (True, False):
        model, sd = build_model(cfg, "cuda:0", 9)
        model.train()
        mask = sd["mask"].cuda().view(1, 1, R, R, R)
        if fused:
            opt = losses.get_optimizer(cfg, model.parameters())
            assert isinstance(opt, train_ops.FusedAdam)
            ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
        else:
            opt = torch.optim.Adam(model.parameters(), lr=cfg.optim.lr, betas=(cfg.optim.beta1, 0.999), eps=cfg.optim.eps)
            ema = _RefEMA(list(model.parameters()), cfg.model.ema_rate)
        state = dict(optimizer=opt, model=model, ema=ema, step=0)
        step_fn = losses.get_step_fn(sde, train=True, optimize_fn=losses.optimization_manager(cfg), mask=mask)
        g = torch.Generator(device="cuda").manual_seed(2)
        batch = torch.randn(B, 4, R, R, R, device="cuda", generator=g).clamp(-1, 1) * mask
        runs.append((state, step_fn, batch))
    (sa, fa, batch), (sb, fb, _) = runs
    for it in range(3):
        torch.manual_seed(100 + it)  # same labels / noise in both runs
        la = fa(sa, batch)["loss"].item()
        torch.manual_seed(100 + it)
        lb = fb(sb, batch)["loss"].item()
        assert abs(la - lb) <= 1e-6 * abs(lb), (it, la, lb)
        with torch.no_grad():
            for (n, p), (_, q) in zip(sa["model"].named_parameters(), sb["model"].named_parameters()):
                assert torch.allclose(p, q, rtol=2e-6, atol=1e-8), (it, n)
                q.copy_(p)
            for s, e in zip(sa["ema"].shadow_params, sb["ema"].shadow_params):
                assert torch.allclose(s, e, rtol=2e-6, atol=1e-8), it
    assert sa["step"] == sb["step"] == 3 and sa["ema"].num_updates == 3
    assert float(sa["optimizer"].state_dict()["state"][2]["step"]) == 3.0


def test_allreduce_grads_over_caller_communicator():
    """mdb_allreduce_grads drives ncclAllReduce on a communicator the HOST created (here a 1-rank communicator made through
    ctypes on the process's libnccl): sum over the ranks, then the mean scaling by the world size the caller states."""
    import ctypes
    from meshdiffusion_b200 import _native
    L = _native.lib()
    nccl = ctypes.CDLL("libnccl.so.2")
    comm = ctypes.c_void_p()
    dev = (ctypes.c_int * 1)(torch.cuda.current_device())
    assert nccl.ncclCommInitAll(ctypes.byref(comm), 1, dev) == 0
    try:
        g = torch.arange(1000, device="cuda", dtype=torch.float32)
        _native.check(L.mdb_allreduce_grads(comm, _native.ptr(g), g.numel(), 1, _native.current_stream()))
        torch.cuda.synchronize()
        assert torch.equal(g, torch.arange(1000, device="cuda", dtype=torch.float32))
        _native.check(L.mdb_allreduce_grads(comm, _native.ptr(g), g.numel(), 2, _native.current_stream()))
        torch.cuda.synchronize()
        assert torch.equal(g, torch.arange(1000, device="cuda", dtype=torch.float32) * 0.5)
    finally:
        nccl.ncclCommDestroy(comm)
