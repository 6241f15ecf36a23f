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


def test_fused_adam_ema_matches_torch():
    from meshdiffusion_b200 import train_ops
    from meshdiffusion_b200.diffusion.models.ema import ExponentialMovingAverage
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(128, 64, 3, 3, 3), (512,), (256, 128), (7,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    opt = torch.optim.Adam(ref_p, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    ema_ref = ExponentialMovingAverage(ref_p, decay=0.9999)
    ema_ours = [p.detach().clone() for p in our_p]
    fused = train_ops.FusedAdamEMA(our_p, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, ema_params=ema_ours)
    n_upd = 0
    for it in range(4):
        grads = [torch.randn(s, device="cuda", generator=g) * (3.0 if it % 2 else 0.01) for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad = gr.clone(); q.grad = gr.clone()
        lr = 2e-5 * min((it + 1) / 5000, 1.0)
        for group in opt.param_groups:
            group["lr"] = lr
        tn_ref = torch.nn.utils.clip_grad_norm_(ref_p, max_norm=1.0)
        opt.step()
        ema_ref.update(ref_p)
        n_upd += 1
        decay = min(0.9999, (1 + n_upd) / (10 + n_upd))
        tn = fused.step(lr=lr, max_norm=1.0, ema_decay=decay)
        assert abs(tn.item() - tn_ref.item()) <= 1e-5 * tn_ref.item()
        for p, q in zip(ref_p, our_p):
            assert torch.allclose(q, p, rtol=2e-6, atol=1e-9)
        for s, e in zip(ema_ref.shadow_params, ema_ours):
            assert torch.allclose(e, s, rtol=2e-6, atol=1e-9)


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
