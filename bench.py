#!/usr/bin/env python
"""Benchmark of the hot path: denoising steps of the res64 unconditional sampler (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo, N ranks (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host cores (oracle port)

A "step" is one denoising step of pc_sampler for one batch: U-Net evaluation + ancestral update over
[batch, 4, 64, 64, 64]. metric = sample-steps/s = batch * steps / time, whole job (sum over ranks). samples/s for the
full 999-evaluation run is value / 999 and is reported as `samples_per_s`.

  value     : device-resident loop (mdb_sampler_run: state, noise and coefficients never leave HBM), CUDA events.
  e2e       : same steps through the public Python API (model(x, labels) + fused update) with the state copied
              host->device from pinned memory before and device->host after EVERY step.
  roofline  : the tcgen05 implicit-GEMM convolution kernel: algorithmic FLOPs / CUDA-event time of exactly those
              launches inside one forward, against the measured sustained bf16 cuBLAS peak (MEASURED_PEAKS.json).
  cpu_baseline : the oracle port of the reference network + update on the host cores, B=1 (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "denoising sample-steps/sec at res-64 (4x64^3), uncond_gen PC sampler"
UNIT = "sample-steps/s"
N_EVALS_FULL = 999  # pc_sampler's unconditional loop evaluates the network N-1 = 999 times (sampling.py:471)


def host_threads():
    """oneDNN conv3d scales to ~32 threads on this host class and gets slower beyond (measured on the 128-core GPU box:
    16 thr 6.7 s, 32 thr 6.1 s, 64 thr 8.3 s, 128 thr 46 s per res64 forward) -- use what is actually useful."""
    return int(os.environ.get("MDB_CPU_THREADS", min(32, os.cpu_count() or 1)))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


def build_model(precision, batch, device, res=64):
    from configs import res64, res128
    from meshdiffusion_b200.diffusion.models import utils as mutils
    from meshdiffusion_b200.diffusion.models.init_utils import random_init_nondegenerate
    cfg = (res128 if res == 128 else res64).get_config()
    cfg.model.compute_dtype = precision
    cfg.model.engine_max_batch = batch
    cfg.device = device
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    random_init_nondegenerate(model.module)
    model.eval()
    return cfg, model


def run_ours(args):
    from meshdiffusion_b200 import _native
    from meshdiffusion_b200.diffusion import sde_lib, sampling
    from meshdiffusion_b200.geometry.dmtet import grid_mask_from_tets
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the single JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    B, R, K, W = args.batch, args.res, args.steps, args.warmup
    cfg, model = build_model(args.precision, B, device, R)
    net = model.module
    sde = sde_lib.VPSDE(cfg.model.beta_min, cfg.model.beta_max, cfg.model.num_scales, device=device)
    mask = grid_mask_from_tets(R).to(device)
    net.mask.data[:] = mask.view(1, 1, R, R, R)
    mask_flat = mask.reshape(-1).contiguous()
    timesteps = torch.linspace(sde.T, 1e-3, sde.N, device=device)
    idx = (timesteps * (sde.N - 1)).long()
    labels_all = (timesteps * (sde.N - 1)).cpu().tolist()
    betas = sde.discrete_betas[idx].cpu().tolist()
    stds = sde.sqrt_1m_alphas_cumprod[idx].cpu().tolist()
    g = torch.Generator(device=device).manual_seed(42 + rank)
    x = (torch.randn(B, 4, R, R, R, device=device, generator=g) * mask).contiguous()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def native_steps(first, n):
        return sampling._native_loop(net, x, mask_flat, labels_all[first:], betas[first:], stds[first:], n, 42 + rank)

    # ---- device-resident loop ("value")
    native_steps(0, W)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native_steps(W, K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * K / (ms * 1e-3)

    # ---- end to end through the public API with host buffers every step ("e2e")
    host_x = torch.empty(B, 4, R, R, R, pin_memory=True)
    host_x.copy_(x.cpu())
    host_out = torch.empty(B, 4, R, R, R, pin_memory=True)
    vec = torch.ones(B, device=device)

    def api_step(i):
        xd = host_x.to(device, non_blocking=True)
        eps = model(xd, vec * labels_all[i])
        xd, x_mean = sampling._fused_update(eps, xd, torch.randn_like(xd), mask_flat, betas[i], stds[i])
        host_out.copy_(x_mean, non_blocking=True)
        host_x.copy_(xd, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    with net.frozen():  # what pc_sampler does around its loop: weights cannot change between steps
        for i in range(W):
            api_step(i)
        barrier()
        e0.record()
        for i in range(W, W + K):
            api_step(i)
        e1.record()
        barrier()
    ms_e2e = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e_value = world * B * K / (ms_e2e * 1e-3)
    state_bytes = B * 4 * R ** 3 * 4

    # ---- roofline of the dominant kernel (rank 0): per-launch CUDA events inside one forward
    roofline, launches_per_forward, info = None, None, net.engine_info()
    if rank == 0:
        labels = vec * labels_all[W]
        prof = net.profile(x, labels)
        if args.dump_profile:
            with open(args.dump_profile, "w") as f:
                json.dump(prof, f)
        launches_per_forward = len(prof) + 2  # + stats memset, + second kernel of the temb step
        conv_ms = sum(t for n, t in prof if _is_conv_gemm(n))
        all_gemm_ms = sum(t for n, t in prof if _is_gemm(n))
        peaks, src = measured_peaks()
        flops = info["flops_per_sample"] * B
        # FLOPs of the non-conv GEMMs (attention, stem) are < 3 % of the total; the roofline is quoted on all GEMM
        # launches of the tcgen05 kernel together: algorithmic FLOPs / their summed duration
        peak = peaks["bf16_tflops_sustained"] * (0.5 if args.precision == "tf32" else 1.0)
        ach = flops / (all_gemm_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv3d / NIN / attention)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": ncu_traffic(),
                    "peak_source": f"{src} bf16_tflops_sustained" + (" x0.5 (tf32 rate)" if args.precision == "tf32" else ""),
                    "gemm_ms_per_forward": all_gemm_ms, "conv_ms_per_forward": conv_ms,
                    "forward_ms": sum(t for _, t in prof), "gemm_launches_per_forward": info["gemm_launches"],
                    "algorithmic_flops_per_launch_avg": flops / info["gemm_launches"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and R == 64:
        cpu = cpu_baseline_port(net, mask, steps=1)

    # ---- the same device-resident loop with TF32 operands (the parity-grade mode: 1.5e-3 rel-L2 vs fp32, the class of
    #      arithmetic the reference's own GPU path uses), reported next to the bf16 headline
    other = None
    if args.precision == "bf16" and not args.no_tf32_leg:
        net.release_engine()
        del model, net
        torch.cuda.empty_cache()
        cfg2, model2 = build_model("tf32", B, device, R)
        net2 = model2.module
        net2.mask.data[:] = mask.view(1, 1, R, R, R)
        x2 = x.clone()
        steps2 = lambda first, n: sampling._native_loop(net2, x2, mask_flat, labels_all[first:], betas[first:], stds[first:], n, 42 + rank)
        steps2(0, W)
        barrier()
        e0.record()
        steps2(W, K)
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms2], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = t.item()
        other = {"dtype": "tf32", "value": world * B * K / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / K}
        net2.release_engine()

    # ---- secondary: one training step (BASELINE configs[2]) through tools/bench_train.py in a fresh process
    train = None
    if rank == 0 and world == 1 and not args.no_train_leg and R == 64:
        import subprocess
        torch.cuda.empty_cache()
        try:
            r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "bench_train.py"),
                                "--batch", "16", "--iters", "1", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
            t = json.loads(r.stdout.strip().splitlines()[-1])
            train = {"value": t["value"], "unit": t["unit"], "workload": t["config"]["workload"], "split_ms_per_step": t["split_ms_per_step"],
                     "fwd_bwd_tflops": t["roofline"]["achieved"], "fwd_bwd_frac_of_peak": t["roofline"]["frac"]}
        except Exception as ex:  # informational leg: never fails the bench line
            train = {"error": str(ex)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC if R == 64 else METRIC.replace("res-64 (4x64^3)", "res-128 (4x128^3)"), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic (random-init non-degenerate weights, N(0,1)*grid_mask state)",
            "config": {"workload": f"res{R}.py uncond_gen, batch={B}/GPU, PC sampler (ancestral_sampling + none), steps of the N=1000 schedule",
                       "batch_per_gpu": B, "image_size": R, "precision": args.precision,
                       "l2_policy": "inputs larger than L2: per-step activation working set is several GB at batch 32"},
            "samples_per_s": value / N_EVALS_FULL,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": 2 * state_bytes,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(K * (launches_per_forward + 2)) if launches_per_forward else None,
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "tf32_operands": other,
            "engine": info, "train": train,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic():
    """dram read+write bytes per launch of the dominant launch type (128->128 conv @64^3) from the committed
    `ncu --set full` capture (profiles/r01_ncu_prof_conv_v2.txt, taken at batch 8), scaled to this run's batch is NOT
    attempted: the figure is reported as captured, with its batch."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_prof_conv_v2.txt")
    try:
        rd = wr = None
        for line in open(path):
            if "dram__bytes_read.sum =" in line and rd is None:
                v, u = line.split("=")[1].split()[:2]; rd = float(v) * (1e9 if u.startswith("G") else 1e6)
            if "dram__bytes_write.sum =" in line and wr is None:
                v, u = line.split("=")[1].split()[:2]; wr = float(v) * (1e9 if u.startswith("G") else 1e6)
        return {"bytes_per_launch": rd + wr, "batch": 8, "launch": "conv3x3x3 128->128 @64^3",
                "algorithmic_bytes_per_launch": 2 * 8 * 64 ** 3 * 128 * 2 + 27 * 128 * 128 * 2}
    except Exception:
        return None


def _is_gemm(name):
    return (".conv" in name or ".nin" in name or name.endswith(".gemm") or name.endswith(".qk") or name.endswith(".pv")
            or name.endswith(".proj") or name.startswith("down"))


def _is_conv_gemm(name):
    return ".conv" in name or name.startswith("down")


def cpu_baseline_port(net, mask, steps=1, threads=None):
    """Oracle port of the reference network + update on the host cores: B=1, `steps` steps after one warm-up."""
    from oracle import sampler_oracle, unet_oracle
    threads = threads or host_threads()
    torch.set_num_threads(threads)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    arch = dict(net.arch)
    sde = sampler_oracle.VPSDETables()
    R = arch["image_size"]
    m = mask.detach().cpu().view(1, R, R, R)
    x = torch.randn(1, 4, R, R, R) * m
    ts = torch.linspace(1, 1e-3, 1000)
    fn = lambda a, t: unet_oracle.unet_forward(sd, arch, a, t)
    with torch.no_grad():
        x, _ = sampler_oracle.ancestral_update(sde, fn, x, ts[0] * torch.ones(1), torch.randn_like)
        t0 = time.perf_counter()
        for i in range(steps):
            x, xm = sampler_oracle.ancestral_update(sde, fn, x * m, ts[1 + i] * torch.ones(1), torch.randn_like)
        dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"B=1, {steps} step(s) of the same res64 sampler after 1 warm-up (oracle port of the reference modules, torch CPU fp32)",
            "s_per_step": dt / steps}


def run_reference(args):
    """Reference arm: the reference's algorithm on the host cores (the reference itself needs CUDA for its SDE
    tables and ships no installable package; the oracle port is its CPU restatement, pinned by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from configs import res64
    from meshdiffusion_b200.diffusion.models import utils as mutils
    from meshdiffusion_b200.geometry.dmtet import grid_mask_from_tets
    from oracle import sampler_oracle, unet_oracle
    from meshdiffusion_b200.diffusion.models.init_utils import random_init_nondegenerate
    cfg = res64.get_config()
    cfg.device = torch.device("cpu")
    torch.manual_seed(0)
    net = mutils.create_model(cfg, use_parallel=False)
    random_init_nondegenerate(net)
    threads = host_threads()
    torch.set_num_threads(threads)
    R = 64
    mask = grid_mask_from_tets(R).view(1, R, R, R)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    sd["mask"] = mask.view(1, 1, R, R, R).clone()
    arch = dict(net.arch)
    sde = sampler_oracle.VPSDETables()
    ts = torch.linspace(1, 1e-3, 1000)
    fn = lambda a, t: unet_oracle.unet_forward(sd, arch, a, t)
    x = torch.randn(1, 4, R, R, R) * mask
    K, W = args.steps, args.warmup
    with torch.no_grad():
        for i in range(W):
            x, _ = sampler_oracle.ancestral_update(sde, fn, x * mask, ts[i] * torch.ones(1), torch.randn_like)
        t0 = time.perf_counter()
        for i in range(W, W + K):
            x, _ = sampler_oracle.ancestral_update(sde, fn, x * mask, ts[i] * torch.ones(1), torch.randn_like)
        dt = time.perf_counter() - t0
    value = K / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "res64.py uncond_gen PC sampler; each step = a bounded sample (1 of the 32 grids of the batch) on the host cores",
                   "batch_per_step": 1, "image_size": R},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "1 grid of the 32-grid batch per step (U-Net evaluation + ancestral update), torch CPU fp32"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="grids per GPU (default 32 = BASELINE configs[1]; 8 with --res 128 = configs[3])")
    ap.add_argument("--res", type=int, default=64, choices=[64, 128], help="64 = the metric's config (default); 128 = BASELINE configs[3] (secondary)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tf32-leg", action="store_true", help="skip the secondary TF32-operand measurement")
    ap.add_argument("--no-train-leg", action="store_true", help="skip the secondary training-step measurement (tools/bench_train.py)")
    ap.add_argument("--dump-profile", default=None, help="write the per-launch CUDA-event times of one forward as JSON")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 8 if args.res == 128 else 32
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args)


if __name__ == "__main__":
    main()
