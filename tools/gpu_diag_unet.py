"""Full score-network forward on the GPU vs the oracle (torch fp32, TF32 off) + first timings. One process per case."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, json, time, torch
sys.path.insert(0, %r)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from configs import res64, res128
from meshdiffusion_b200.diffusion.models import utils as mutils
from oracle import unet_oracle
name, B, prec, do_oracle, do_profile = json.loads(sys.argv[1])
cfg = res128.get_config() if name == "res128" else res64.get_config()
if name == "tiny":
    cfg.data.image_size = 16; cfg.model.nf = 32; cfg.model.ch_mult = (1, 2); cfg.model.num_res_blocks = 1; cfg.model.attn_resolutions = (8,)
if name == "mid":
    cfg.data.image_size = 32; cfg.model.nf = 64; cfg.model.ch_mult = (1, 2, 2, 2); cfg.model.num_res_blocks = 1; cfg.model.attn_resolutions = (8,)
cfg.model.compute_dtype = prec
cfg.device = torch.device("cuda:0")
torch.manual_seed(0)
model = mutils.create_model(cfg)
net = model.module
sd = unet_oracle.nondegenerate_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, seed=1)
R = cfg.data.image_size
g = torch.Generator().manual_seed(5)
sd["mask"] = (torch.rand(1, 1, R, R, R, generator=g) < 0.3).float()
net.load_state_dict(sd)
x = (torch.randn(B, 4, R, R, R, generator=g) * sd["mask"]).cuda()
labels = (torch.rand(B, generator=g) * 999).cuda()
t0 = time.time()
out = model(x, labels); torch.cuda.synchronize()
t_first = time.time() - t0
res = {"first_call_s": t_first, "info": net.engine_info(), "finite": bool(torch.isfinite(out).all()), "out_absmax": out.abs().max().item()}
if do_oracle:
    sdg = {k: v.cuda() for k, v in sd.items()}
    arch = unet_oracle.arch_from_config(cfg)
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sdg, arch, x, labels)
    d = (out - ref).abs()
    res["max_err_over_max"] = d.max().item() / ref.abs().max().item()
    res["rel_l2"] = (d.pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
    res["ref_absmax"] = ref.abs().max().item()
    # per-element rtol-style check: |d| <= atol + rtol*|ref|
    res["frac_within_1e-3"] = ((d <= 1e-3 * ref.abs() + 1e-3 * ref.abs().max()).float().mean()).item()
# timing
for _ in range(2): model(x, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n): model(x, labels)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
res["ms_per_forward"] = ms
res["tflops"] = res["info"]["flops_per_sample"] * B / (ms * 1e-3) / 1e12
if do_profile:
    prof = net.profile(x, labels)
    agg = {}
    for nme, t in prof:
        key = nme.split(":")[0] if ":" in nme else nme
        agg[nme] = t
    top = sorted(prof, key=lambda p: -p[1])[:25]
    res["profile_total_ms"] = sum(t for _, t in prof)
    res["profile_top"] = top
    kinds = {}
    for nme, t in prof:
        k = "gemm" if (".conv" in nme or ".nin" in nme or "gemm" in nme or ".qk" in nme or ".pv" in nme or nme.startswith("down")) else nme.split(":")[0].split(".")[-1]
        kinds[k] = kinds.get(k, 0) + t
    res["profile_kinds"] = kinds
print("RESULT", json.dumps(res))
''' % ROOT

cases = [
    ("tiny", 2, "tf32", 1, 0), ("tiny", 2, "bf16", 1, 0),
    ("mid", 2, "tf32", 1, 0), ("mid", 3, "bf16", 1, 0),
    ("res64", 1, "tf32", 1, 1), ("res64", 2, "bf16", 1, 1),
    ("res64", 8, "bf16", 0, 1), ("res64", 8, "tf32", 0, 0),
    ("res128", 1, "bf16", 1, 1),
]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for c in cases:
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(list(c))], capture_output=True, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        if lines:
            print(c, lines[0][7:], flush=True)
        else:
            print(c, "FAILED rc", r.returncode, (r.stdout[-800:] + r.stderr[-2500:]).replace("\n", " | "), flush=True)
    except subprocess.TimeoutExpired:
        print(c, "TIMEOUT", flush=True)
