"""Python face of the optimiser-side training kernels: forward perturbation + masked DDPM loss (lib/diffusion/losses.py:54-85
of the reference), global-norm clipping + Adam (losses.py:26-52) and the EMA update (models/ema.py:43-64).

`FusedAdam` IS a `torch.optim.Adam` (same constructor, `param_groups`, per-parameter `state` entries `step` / `exp_avg` /
`exp_avg_sq`, hence the same `state_dict()` -- checkpoints interchange with the reference's); only `step()` differs: one
native multi-tensor pass applies the clip coefficient, the Adam update and, when an EMA is handed in, its update too
(the reference makes ~6 passes over the 364 M parameters). There is no CPU path: stepping CPU parameters raises.
"""
import numpy as np
import torch

from . import _native


def chunk_table(numels, device):
    """int32 [n_chunks, 2] (tensor index, chunk index) covering every tensor in mdb_chunk_elems()-element pieces."""
    ch = _native.lib().mdb_chunk_elems()
    counts = [(int(n) + ch - 1) // ch for n in numels]
    t = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
    c = np.concatenate([np.arange(k, dtype=np.int32) for k in counts]) if counts else np.zeros(0, np.int32)
    return torch.from_numpy(np.stack([t, c], axis=1).copy()).to(device)


class _Tables:
    """Device pointer tables of parallel tensor lists (params, grads, ...) + sizes + chunk table, rebuilt only when a
    tensor moved (the keys are the data pointers themselves)."""

    def __init__(self):
        self.key, self.ptrs, self.numels, self.chunks, self.scratch = None, None, None, None, None

    def get(self, lists):
        key = tuple(t.data_ptr() for lst in lists for t in lst)
        if key != self.key:
            dev = lists[0][0].device
            n = len(lists[0])
            self.ptrs = [torch.tensor(key[i * n:(i + 1) * n], dtype=torch.int64, device=dev) for i in range(len(lists))]
            numels = [t.numel() for t in lists[0]]
            self.numels = torch.tensor(numels, dtype=torch.int64, device=dev)
            self.chunks = chunk_table(numels, dev)
            self.scratch = torch.zeros(self.chunks.shape[0], dtype=torch.float64, device=dev)
            self.key = key
        return self


def _check_cuda(tensors, what):
    for t in tensors:
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise _native.NativeError(f"{what}: the native optimiser kernels take contiguous fp32 CUDA tensors; there is no CPU path")


def ddpm_perturb(batch, noise, mask, sqrt_ac, sqrt_1mac):
    """x_t = (sqrt(abar_t) x_0 + sqrt(1 - abar_t) eps) * mask, one pass (losses.py:63-66); per-sample coefficients [B]."""
    L = _native.lib()
    B, C = batch.shape[0], batch.shape[1]
    V = batch[0, 0].numel()
    batch, noise = batch.float().contiguous(), noise.float().contiguous()
    m = mask.reshape(-1).float().contiguous()
    out = torch.empty_like(batch)
    _native.check(L.mdb_ddpm_perturb(_native.ptr(batch), _native.ptr(noise), _native.ptr(m), _native.ptr(sqrt_ac.float().contiguous()),
                                     _native.ptr(sqrt_1mac.float().contiguous()), _native.ptr(out), B, C, V, _native.current_stream()))
    return out


def ddpm_loss(pred, noise, mask, want_grad=False, mask_sum=None):
    """losses.py:69-78 on fp32 [B,C,R,R,R] tensors with mask broadcastable [1,1,R,R,R]; returns (loss, grad or None)."""
    L = _native.lib()
    B, C = pred.shape[0], pred.shape[1]
    V = pred[0, 0].numel()
    m = mask.reshape(-1).float().contiguous()
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    scratch = torch.empty(1, device=pred.device, dtype=torch.float64)
    if mask_sum is None:
        mask_sum = float(m.sum().item())
    _native.check(L.mdb_ddpm_loss(_native.ptr(pred.contiguous()), _native.ptr(noise.contiguous()), _native.ptr(m),
                                  mask_sum, _native.ptr(loss), _native.ptr(grad), _native.ptr(scratch), B, C, V,
                                  _native.current_stream()))
    return loss, grad


class DDPMLossFn(torch.autograd.Function):
    """The masked l2 DDPM loss as one autograd node: forward computes the loss AND d loss / d pred in one pass over the
    prediction, so `loss.backward()` of the reference's step_fn (losses.py:124) hands the network a ready gradient."""

    @staticmethod
    def forward(ctx, pred, noise, mask, mask_sum):
        loss, grad = ddpm_loss(pred.detach().float(), noise.float(), mask, want_grad=True, mask_sum=mask_sum)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None, None


def ema_update(shadow, params, decay, tables=None):
    """shadow -= (1 - decay) (shadow - p), one native pass over all tensors."""
    L = _native.lib()
    _check_cuda(shadow, "EMA shadow")
    _check_cuda([p.data for p in params], "EMA parameters")
    t = (tables or _Tables()).get([shadow, [p.data for p in params]])
    _native.check(L.mdb_ema_update(_native.ptr(t.ptrs[0]), _native.ptr(t.ptrs[1]), _native.ptr(t.numels), _native.ptr(t.chunks),
                                   t.chunks.shape[0], float(decay), _native.current_stream()))


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=False)
        self._tables = {}
        self._coef = None
        self._norm = None

    def grad_norm_coef(self, max_norm):
        """clip_grad_norm_'s coefficient min(1, max_norm / (||g|| + 1e-6)) over every gradient of the optimiser, kept on the
        device for the next step(); returns the total norm (device scalar). The gradients are not rescaled in place: step()
        applies the coefficient on the fly."""
        L = _native.lib()
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not grads:
            return None
        _check_cuda(grads, "gradients")
        t = self._tables.setdefault("norm", _Tables()).get([grads])
        if self._coef is None or self._coef.device != grads[0].device:
            self._coef = torch.ones(1, device=grads[0].device)
            self._norm = torch.zeros(1, device=grads[0].device)
        _native.check(L.mdb_grad_clip_coef(_native.ptr(t.ptrs[0]), _native.ptr(t.numels), _native.ptr(t.chunks), t.chunks.shape[0],
                                           float(max_norm), _native.ptr(self._coef), _native.ptr(self._norm), _native.ptr(t.scratch),
                                           _native.current_stream()))
        self._clip_pending = True
        return self._norm

    @torch.no_grad()
    def step(self, closure=None, ema=None):
        """One Adam update of every parameter that has a gradient. `ema`: an ExponentialMovingAverage whose update is folded
        into the same pass (returns True when it was; the caller then must not call ema.update() for this step)."""
        if closure is not None:
            raise NotImplementedError("FusedAdam.step does not take a closure")
        L = _native.lib()
        coef = self._coef if getattr(self, "_clip_pending", False) else None
        self._clip_pending = False
        ema_done = False
        single_group = len(self.param_groups) == 1
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("FusedAdam: amsgrad / maximize are not supported")
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            _check_cuda([p.data for p in params], "parameters")
            _check_cuda([p.grad for p in params], "gradients")
            for p in params:
                st = self.state[p]
                if len(st) == 0:  # torch.optim.Adam's lazy state, same entries and dtypes
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = [self.state[p]["step"] for p in params]
            torch._foreach_add_(steps, 1.0)
            step = int(steps[0].item())
            lists = [[p.data for p in params], [p.grad for p in params], [self.state[p]["exp_avg"] for p in params],
                     [self.state[p]["exp_avg_sq"] for p in params]]
            _check_cuda(lists[2] + lists[3], "Adam state")
            shadow, decay = None, 0.0
            if ema is not None and single_group:
                trainable = [p for p in group["params"] if p.requires_grad]
                if len(trainable) == len(params) and len(ema.shadow_params) == len(params) and all(s.is_cuda for s in ema.shadow_params):
                    shadow = ema.shadow_params
                    _check_cuda(shadow, "EMA shadow")
                    decay = ema._current_decay()
                    lists.append(shadow)
                    ema_done = True
            t = self._tables.setdefault(("adam", gi, shadow is not None), _Tables()).get(lists)
            beta1, beta2 = group["betas"]
            _native.check(L.mdb_adam_ema_step(_native.ptr(t.ptrs[0]), _native.ptr(t.ptrs[1]), _native.ptr(t.ptrs[2]), _native.ptr(t.ptrs[3]),
                                              _native.ptr(t.ptrs[4]) if shadow is not None else None, _native.ptr(t.numels),
                                              _native.ptr(t.chunks), t.chunks.shape[0], float(group["lr"]), beta1, beta2, group["eps"],
                                              float(group["weight_decay"]), step, _native.ptr(coef), float(decay), _native.current_stream()))
        return ema_done
