"""Two-GPU tests (skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the data-parallel training exchange -- gradient buckets all-reduced over NCCL on a side stream while the backward pass is
still running -- against the single-rank run on the concatenated batch, and against the blocking whole-buffer all-reduce."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu

CHILD = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from helpers import build_model, tiny_config
from meshdiffusion_b200 import train_ops
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
cfg = tiny_config("res64", "bf16")
cfg.model.dropout = 0.0
cfg.model.grad_bucket_mb = 1          # several buckets even for the tiny network
model, sd = build_model(cfg, f"cuda:{rank}", 9)
model.train()
net = model.module
R, B = 16, 2
mask = sd["mask"].to(dev).view(1, 1, R, R, R)
g = torch.Generator().manual_seed(5)
x_all = (torch.randn(world * B, 4, R, R, R, generator=g) * sd["mask"].view(1, 1, R, R, R)).to(dev)
noise_all = torch.randn(world * B, 4, R, R, R, generator=g).to(dev)
labels_all = torch.randint(0, 1000, (world * B,), generator=g).to(dev)
msum = float(mask.sum().item())

def grads(x, noise, labels, overlap):
    for p in net.parameters():
        p.grad = None
    net.grad_overlap = overlap
    net.reduce_in_backward = True
    loss = train_ops.DDPMLossFn.apply(model(x, labels), noise, mask, msum)
    loss.backward()
    pending = net._pending_reduce is not None
    net.allreduce_grads()
    torch.cuda.synchronize()
    return net._flat_grad.clone(), pending

sl = slice(rank * B, (rank + 1) * B)
g_overlap, was_pending = grads(x_all[sl], noise_all[sl], labels_all[sl], True)
assert was_pending, "the backward pass did not launch the bucketed reductions"
assert len(net._grad_buckets()) >= 3, net._grad_buckets()
g_block, was_pending2 = grads(x_all[sl], noise_all[sl], labels_all[sl], False)
assert not was_pending2
assert torch.allclose(g_overlap, g_block, rtol=1e-6, atol=1e-9), "bucketed/overlapped and blocking all-reduce disagree"
# every rank holds the same averaged gradient
other = g_overlap.clone()
dist.broadcast(other, src=0)
assert torch.equal(other, g_overlap)
if rank == 0:
    # single-rank reference: the concatenated batch, no exchange (the mean over 2B grids == the mean of the two rank means)
    dist_world = dist.get_world_size
    net.reduce_in_backward = False
    for p in net.parameters():
        p.grad = None
    loss = train_ops.DDPMLossFn.apply(model(x_all, labels_all), noise_all, mask, msum)
    loss.backward()
    torch.cuda.synchronize()
    ref = net._flat_grad
    err = ((g_overlap - ref).double().norm() / ref.double().norm()).item()
    print(f"DP_GRAD_REL_L2 {err:.3e} buckets {len(net._grad_buckets())}")
    assert err < 1e-3, err  # bf16 operands: the two runs round different partial sums
dist.barrier()
print("DP_RANK_OK", rank)
dist.destroy_process_group()
''' % (ROOT, ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_overlapped_allreduce_matches_single_rank_double_batch(tmp_path):
    script = os.path.join(tmp_path, "dp_child.py")
    open(script, "w").write(CHILD)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.count("DP_RANK_OK") == 2 and "DP_GRAD_REL_L2" in r.stdout
