"""Bring-up diagnostic for the tcgen05 weight-gradient kernel: one conv case under both descriptor conventions."""
import os
import subprocess
import sys

CODE = r'''
import torch, torch.nn.functional as F, sys
sys.path.insert(0, ".")
from meshdiffusion_b200 import ops
torch.backends.cudnn.allow_tf32 = False
for (B, Cin, Cout, R, k, s) in [(2, 64, 128, 16, 3, 1), (3, 96, 64, 4, 1, 1), (3, 128, 256, 8, 3, 1)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, Cin, R, R, R, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, k, device="cuda", generator=g) * 0.05).requires_grad_(True)
    y = F.conv3d(x, w, None, padding=k // 2)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16().float()
    y.backward(dy)
    nd = lambda t: t.permute(0, 2, 3, 4, 1).contiguous().bfloat16()
    try:
        dw, _ = ops.conv3d_backward(nd(dy), nd(x.detach()), w.detach(), stride=s, want_dx=False)
        err = (dw - w.grad).abs().max().item() / w.grad.abs().max().item()
        print(f"  case B{B} {Cin}->{Cout} R{R} k{k}: rel max err {err:.3e}  (|ref| {w.grad.abs().max().item():.3f}, |got| {dw.abs().max().item():.3f})")
    except Exception as e:
        print("  case failed:", e)
        break
'''
for dbg in ("0", "1"):
    print(f"MDB_WG_DBG={dbg}")
    env = dict(os.environ, MDB_WG_DBG=dbg)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    if r.returncode:
        print("  exit", r.returncode, r.stderr[-1500:])
