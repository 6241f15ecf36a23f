"""Diagnostic: full-size res64 Adam steps on the engine, reporting the first non-finite gradient tensor per step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_model, ddpm_loss, full_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
cfg = full_config("res64", "bf16"); cfg.model.dropout = 0.0
model, sd = build_model(cfg, "cuda:0", 5)
net = model.module; net.train()
R = 64
mask = sd["mask"].cuda().view(1, 1, R, R, R)
params = [p for p in net.parameters() if p.requires_grad]
names = [n for n, p in net.named_parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=lr)
g = torch.Generator(device="cuda").manual_seed(4)
data = (torch.rand(B, 4, R, R, R, device="cuda", generator=g) * 2 - 1) * mask
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 8
for it in range(STEPS):
    labels = torch.randint(0, 1000, (B,), device="cuda", generator=g).float()
    noise = torch.randn(data.shape, device="cuda", generator=g)
    x = (0.7 * data + 0.7 * noise) * mask
    opt.zero_grad()
    pred = model(x, labels)
    lo = ddpm_loss(pred, noise, mask)
    import ctypes, numpy as np
    from meshdiffusion_b200 import _native
    L = _native.lib(); cnt = ctypes.c_longlong()
    torch.cuda.synchronize()
    _native.check(L.mdb_unet_debug_stats(net._train_handle, None, 0, ctypes.byref(cnt)))
    buf = np.zeros(cnt.value, dtype=np.int64)
    _native.check(L.mdb_unet_debug_stats(net._train_handle, buf.ctypes.data_as(ctypes.c_void_p), cnt.value, ctypes.byref(cnt)))
    w = buf.reshape(-1, 4).astype(np.float64)
    st = np.stack([w[:, 0] / 2 ** 24 + w[:, 1] * 65536.0, w[:, 2] / 2 ** 24 + w[:, 3] * 65536.0], 1)
    neg = int((st[:, 1] < 0).sum()); mx = st[:, 1].max(); amax = np.abs(st[:, 0]).max()
    first_neg = int(np.argmax(st[:, 1] < 0)) if neg else -1
    print(f"   stats: {st.shape[0]} (b,c) entries, negative sumsq {neg} (first at {first_neg}), max sumsq {mx:.3e} (int64 limit 5.5e11), max |sum| {amax:.3e}", flush=True)
    lo.backward()
    bad = [(n, int((~torch.isfinite(p.grad)).sum())) for n, p in zip(names, params) if not torch.isfinite(p.grad).all()]
    gn = torch.nn.utils.clip_grad_norm_(params, 1.0)
    big = sorted(((p.grad.abs().max().item(), n) for n, p in zip(names, params)), reverse=True)[:3]
    print(f"step {it}: loss {lo.item():.4f} pred finite {bool(torch.isfinite(pred).all())} |pred| {pred.abs().max().item():.3e} grad norm {gn.item():.4e} nonfinite {bad[:4]} largest {[(round(v,4), n) for v, n in big]}", flush=True)
    if bad or not torch.isfinite(lo):
        break
    opt.step()
    pb = [n for n, p in zip(names, params) if not torch.isfinite(p).all()]
    if pb:
        print("  non-finite parameters after step:", pb[:5]); break
