"""Training-step throughput of the sm_100a engine (BASELINE.json configs[2]: res64 train, synthetic 4x64^3 grids,
bf16 operands, fp32 master weights + Adam + EMA, data-parallel gradient all-reduce).

    python tools/bench_train.py [--batch 16] [--iters 4] [--steps 3] [--warmup 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py ...

One optimiser step = `iters` micro-batches of `batch` grids per GPU (forward + loss + backward each, gradients
accumulated in the flat fp32 buffer), one all-reduce of that buffer, one fused clip + Adam + EMA pass. Prints ONE JSON
line: samples/s over all ranks, the device-time split, and the achieved tensor-core rate (3 x forward FLOPs per sample,
SURVEY 8d) against the measured sustained bf16 peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="micro-batch per GPU")
    ap.add_argument("--iters", type=int, default=4, help="micro-batches per optimiser step (batch*iters = 64/GPU in BASELINE)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--config", default="res64", choices=["res64", "res128", "tiny"])
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--profile", default=None, help="write per-launch device times of one forward+backward as JSON")
    args = ap.parse_args()

    import torch.distributed as dist
    from configs import res64, res128
    from meshdiffusion_b200 import train_ops
    from meshdiffusion_b200.diffusion import sde_lib
    from meshdiffusion_b200.diffusion.models import utils as mutils
    from meshdiffusion_b200.diffusion.evaler import load_grid_mask

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")

    cfg = (res128 if args.config == "res128" else res64).get_config()
    if args.config == "tiny":  # test-size architecture (every layer type, seconds to build)
        cfg.data.image_size, cfg.model.nf, cfg.model.ch_mult = 16, 32, (1, 2)
        cfg.model.num_res_blocks, cfg.model.attn_resolutions = 1, (8,)
    cfg.model.compute_dtype = "bf16"
    cfg.model.dropout = args.dropout
    cfg.device = dev
    torch.manual_seed(42)
    model = mutils.create_model(cfg)
    net = model.module
    net.train()
    R, B = cfg.data.image_size, args.batch
    mask = (load_grid_mask(R, dev) if R in (64, 128) else torch.ones(R, R, R, device=dev)).view(1, 1, R, R, R)
    net.mask.data[:] = mask
    # non-degenerate weights (the reference zero-inits Conv_1 / NIN_3 / head, which would make most gradients vanish)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("Conv_1.weight") or n.endswith("NIN_3.W") or (n.startswith("all_modules.") and p.dim() == 5 and p.shape[0] == 4):
                fan = p[0].numel() if p.dim() > 1 else 1
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (3.0 / fan) ** 0.5)
    sde = sde_lib.VPSDE(cfg.model.beta_min, cfg.model.beta_max, cfg.model.num_scales, device=dev)
    params = [p for p in net.parameters() if p.requires_grad]
    ema = [p.detach().clone() for p in params]
    opt = train_ops.FusedAdamEMA(params, lr=cfg.optim.lr, betas=(cfg.optim.beta1, 0.999), eps=cfg.optim.eps, ema_params=ema)
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    data = [(torch.rand(B, 4, R, R, R, device=dev, generator=gen) * 2 - 1) * mask for _ in range(2)]

    ev = lambda: torch.cuda.Event(enable_timing=True)
    split = {"fwd": 0.0, "loss": 0.0, "bwd": 0.0, "allreduce": 0.0, "optimizer": 0.0, "weight_sync": 0.0}

    def one_step(step, timed):
        marks = []
        for p in params:
            p.grad = None
        last_loss = None
        for it in range(args.iters):
            batch = data[(step + it) % 2]
            labels = torch.randint(0, sde.N, (B,), device=dev, generator=gen)
            noise = torch.randn(batch.shape, device=dev, generator=gen)
            a = sde.sqrt_alphas_cumprod.to(dev)[labels, None, None, None, None]
            s = sde.sqrt_1m_alphas_cumprod.to(dev)[labels, None, None, None, None]
            x = ((a * batch + s * noise) * mask).contiguous()
            lab = labels.float()
            e0, e1, e2, e3, e4 = ev(), ev(), ev(), ev(), ev()
            e0.record()
            net._train_synced = net._push_parameters(net._train_handle, net._train_synced) if net._train_handle is not None else None
            e1.record()
            pred = net._train_forward(x, lab)
            e2.record()
            loss, dpred = train_ops.ddpm_loss(pred, noise, mask, want_grad=True)
            e3.record()
            none = params[0].grad is None
            net._train_backward(x, lab, dpred)
            e4.record()
            last_loss = loss
            marks.append((e0, e1, e2, e3, e4))
        e5, e6, e7 = ev(), ev(), ev()
        e5.record()
        net.allreduce_grads()
        e6.record()
        lr = cfg.optim.lr * min((step + 1) / cfg.optim.warmup, 1.0) if cfg.optim.warmup > 0 else cfg.optim.lr
        n_upd = step + 1
        opt.step(lr=lr, max_norm=cfg.optim.grad_clip, ema_decay=min(cfg.model.ema_rate, (1 + n_upd) / (10 + n_upd)))
        e7.record()
        if timed:
            torch.cuda.synchronize()
            for (a0, a1, a2, a3, a4) in marks:
                split["weight_sync"] += a0.elapsed_time(a1); split["fwd"] += a1.elapsed_time(a2)
                split["loss"] += a2.elapsed_time(a3); split["bwd"] += a3.elapsed_time(a4)
            split["allreduce"] += e5.elapsed_time(e6); split["optimizer"] += e6.elapsed_time(e7)
        return last_loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for w in range(args.warmup):
        losses.append(one_step(w, False).item())
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for s_ in range(args.steps):
        losses.append(one_step(args.warmup + s_, True).item())
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    import ctypes
    from meshdiffusion_b200 import _native
    L = _native.lib()
    fl, bf, nb, numel = ctypes.c_double(), ctypes.c_double(), ctypes.c_int(), ctypes.c_longlong()
    _native.check(L.mdb_unet_info(net._train_handle, ctypes.byref(fl), None, None, None))
    _native.check(L.mdb_unet_train_info(net._train_handle, ctypes.byref(bf), ctypes.byref(nb), ctypes.byref(numel)))
    samples = args.steps * args.iters * B * world
    peak = 1418.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    tc_ms = split["fwd"] + split["bwd"]
    achieved = (fl.value + bf.value) * args.steps * args.iters * B / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None
    if args.profile and rank == 0:
        x = data[0]
        lab = torch.full((B,), 500.0, device=dev)
        prof_f = net.profile(x, lab) if False else None  # forward profile is exposed through bench.py --dump-profile
        names = ctypes.create_string_buffer(1 << 18)
        msb = (ctypes.c_float * 4096)()
        n = ctypes.c_int()
        pred = net._train_forward(x, lab)
        _native.check(L.mdb_unet_profile_backward(net._train_handle, _native.ptr(pred), _native.ptr(net._flat_grad), B,
                                                  _native.current_stream(), names, len(names), msb, 4096, ctypes.byref(n)))
        rows = list(zip(names.value.decode().strip().split("\n"), [msb[i] for i in range(n.value)]))
        json.dump(rows, open(args.profile, "w"))
    if rank == 0:
        print(json.dumps({
            "metric": "training samples/s (res64 4x64^3 grids, bf16 operands, fp32 master/Adam/EMA)", "value": samples / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.config}.py train, micro-batch {B} x {args.iters} per GPU, dropout {args.dropout}, clip 1.0, Adam + EMA",
                       "global_batch": B * args.iters * world},
            "split_ms_per_step": {k: v / args.steps for k, v in split.items()},
            "flops_per_sample": {"forward": fl.value, "backward": bf.value},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved else None,
                         "note": "forward + backward GEMM FLOPs / (fwd + bwd device time)"},
            "bwd_launches": nb.value, "params": numel.value, "losses": losses,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
