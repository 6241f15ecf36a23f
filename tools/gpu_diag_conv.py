"""Runs each conv parity case in its own process (a trapped kernel kills the CUDA context) and prints a table."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    (1, 128, 128, 16, 1, 1), (2, 384, 128, 16, 1, 1),
    (1, 128, 128, 4, 3, 1), (3, 512, 512, 4, 3, 1), (2, 128, 256, 8, 3, 1),
    (2, 128, 128, 16, 3, 1), (1, 256, 128, 16, 3, 1), (1, 128, 128, 32, 3, 1),
    (2, 128, 128, 16, 3, 2), (2, 256, 256, 8, 3, 2),
    (1, 128, 4, 16, 3, 1), (1, 128, 4, 32, 5, 1),
]

CHILD = r'''
import sys, json, torch, torch.nn.functional as F
sys.path.insert(0, %r)
from meshdiffusion_b200 import ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
B, Cin, Cout, R, k, stride, precision = json.loads(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(B, Cin, R, R, R, device="cuda", generator=g)
w = torch.randn(Cout, Cin, k, k, k, device="cuda", generator=g) / (Cin * k ** 3) ** 0.5
b = torch.randn(Cout, device="cuda", generator=g)
ref = F.conv3d(x, w, b, padding=k // 2) if stride == 1 else F.conv3d(F.pad(x, (0, 1, 0, 1, 0, 1)), w, b, stride=2)
y, stats = ops.conv3d(ops.to_ndhwc(x, precision), w, b, stride=stride, want_stats=True, precision=precision)
torch.cuda.synchronize()
out = ops.from_ndhwc(y)
err = (out - ref).abs().max().item() / ref.abs().max().item()
s_ref = ref.double().sum(dim=(2, 3, 4))
serr = (stats[..., 0] - s_ref).abs().max().item() / s_ref.abs().max().item()
q_ref = (ref.double() ** 2).sum(dim=(2, 3, 4))
qerr = (stats[..., 1] - q_ref).abs().max().item() / q_ref.abs().max().item()
# where is the error?
d = (out - ref).abs()
idx = (d == d.max()).nonzero()[0].tolist()
print("RESULT", json.dumps({"err": err, "sum_err": serr, "sq_err": qerr, "argmax": idx, "out_absmax": out.abs().max().item(), "ref_absmax": ref.abs().max().item()}))
''' % ROOT

for prec in ["bf16", "tf32"]:
    for c in CASES:
        arg = json.dumps(list(c) + [prec])
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, arg], capture_output=True, text=True, timeout=180)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            if lines:
                print(prec, c, lines[0][7:], flush=True)
            else:
                print(prec, c, "FAILED rc", r.returncode, (r.stdout[-600:] + r.stderr[-1200:]).replace("\n", " | "), flush=True)
        except subprocess.TimeoutExpired:
            print(prec, c, "TIMEOUT", flush=True)
