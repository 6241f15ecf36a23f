"""Python face of the optimiser-side training kernels (masked DDPM loss, global-norm clip, fused Adam + EMA)."""
import torch

from . import _native


def ddpm_loss(pred, noise, mask, want_grad=False):
    """losses.py:69-78 on fp32 [B,C,R,R,R] tensors with mask broadcastable [1,1,R,R,R]; returns (loss, grad or None)."""
    L = _native.lib()
    B, C = pred.shape[0], pred.shape[1]
    V = pred[0, 0].numel()
    m = mask.reshape(-1).float().contiguous()
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    scratch = torch.empty(1, device=pred.device, dtype=torch.float64)
    _native.check(L.mdb_ddpm_loss(_native.ptr(pred.contiguous()), _native.ptr(noise.contiguous()), _native.ptr(m),
                                  float(m.sum().item()), _native.ptr(loss), _native.ptr(grad), _native.ptr(scratch), B, C, V,
                                  _native.current_stream()))
    return loss, grad


class FusedAdamEMA:
    """clip_grad_norm_(max_norm) + Adam.step + EMA.update in two kernels over all parameters
    (one norm reduction, one read-modify-write pass). State layout matches torch.optim.Adam's."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, ema_params=None, ema_decay=0.9999):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.ema_decay = lr, betas, eps, ema_decay
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.ema = ema_params
        self.step_count = 0
        dev = self.params[0].device
        self._numels = torch.tensor([p.numel() for p in self.params], dtype=torch.int64, device=dev)
        self._coef = torch.ones(1, device=dev)
        self._norm = torch.zeros(1, device=dev)
        self._scratch = torch.zeros(1, device=dev, dtype=torch.float64)

    def _table(self, tensors):
        return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=self._numels.device)

    def step(self, lr=None, max_norm=None, ema_decay=None):
        L = _native.lib()
        self.step_count += 1
        stream = _native.current_stream()
        grads = self._table([p.grad for p in self.params])
        n = len(self.params)
        coef = None
        if max_norm is not None and max_norm >= 0:
            _native.check(L.mdb_grad_clip_coef(_native.ptr(grads), _native.ptr(self._numels), n, float(max_norm),
                                               _native.ptr(self._coef), _native.ptr(self._norm), _native.ptr(self._scratch), stream))
            coef = self._coef
        # the pointer tables must stay alive until the kernel has been enqueued (torch's allocator would otherwise hand
        # the same block to the next table)
        t_ema = self._table(self.ema) if self.ema is not None else None
        t_p, t_m, t_v = self._table(self.params), self._table(self.exp_avg), self._table(self.exp_avg_sq)
        _native.check(L.mdb_adam_ema_step(_native.ptr(t_p), _native.ptr(grads), _native.ptr(t_m), _native.ptr(t_v),
                                          _native.ptr(t_ema), _native.ptr(self._numels), n,
                                          float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.eps,
                                          self.step_count, _native.ptr(coef),
                                          float(self.ema_decay if ema_decay is None else ema_decay), stream))
        self._keepalive = (t_p, t_m, t_v, t_ema, grads)
        return self._norm
