"""How many host threads should the CPU reference arm use on the GPU box? Times one oracle forward (B=1, res64) per setting."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from configs import res64
from meshdiffusion_b200.diffusion.models import utils as mutils
from oracle import unet_oracle
cfg = res64.get_config(); cfg.device = torch.device("cpu")
net = mutils.create_model(cfg, use_parallel=False)
sd = {k: v.detach() for k, v in net.state_dict().items()}
arch = dict(net.arch)
x = torch.randn(1, 4, 64, 64, 64); t = torch.tensor([500.0])
out = {"cpu_count": os.cpu_count()}
for n in (16, 32, 64):
    torch.set_num_threads(n)
    with torch.no_grad():
        t0 = time.perf_counter(); unet_oracle.unet_forward(sd, arch, x, t); dt = time.perf_counter() - t0
    out[str(n)] = dt
    print(json.dumps(out), flush=True)
