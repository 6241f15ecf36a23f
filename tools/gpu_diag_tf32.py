"""Calibrates the TF32 error of the full res64 network: ours (with / without rounding stored activations) vs the oracle
in true fp32, next to stock torch with TF32 convs (the reference's default GPU arithmetic)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from helpers import build_model, full_config, rel_l2, rel_max
from oracle import synth, unet_oracle
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
cfg = full_config("res64", "tf32")
model, sd = build_model(cfg, "cuda:0", 5)
x, labels = synth.synthetic_inputs(64, 1, 6, sd["mask"]); x, labels = x.cuda(), labels.cuda()
out = model(x, labels)
sdg = {k: v.cuda() for k, v in sd.items()}
with torch.no_grad():
    ref = unet_oracle.unet_forward(sdg, unet_oracle.arch_from_config(cfg), x, labels)
print("RESULT", json.dumps({"max": rel_max(out, ref), "l2": rel_l2(out, ref)}))
''' % (ROOT, ROOT)
for flag in ("0", "1"):
    env = dict(os.environ, MDB_TF32_ROUND_STORE=flag)
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print("round_store", flag, lines[0] if lines else ("FAILED " + r.stderr[-1500:]), flush=True)
