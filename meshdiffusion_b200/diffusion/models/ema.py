"""Exponential moving average of the trainable parameters, API- and checkpoint-compatible with the reference
(lib/diffusion/models/ema.py:10-98): the shadow list is POSITIONAL (`parameters()` order, trainable tensors only), the
effective decay warms up as `min(decay, (1 + n) / (10 + n))`, the update is `shadow -= (1 - decay_n) * (shadow - p)`, and
`state_dict()` is `{'decay', 'num_updates', 'shadow_params'}`.

On the GPU the update is ONE native multi-tensor pass (mdb_ema_update), or no pass of its own at all when the optimiser
folds it into the Adam kernel (train_ops.FusedAdam.step(ema=...), which is what the training step does). Shadow tensors
that live on the CPU (checkpoint tooling) are updated with torch ops.
"""
import torch


def _trainable(parameters):
    return [p for p in parameters if p.requires_grad]


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if not 0.0 <= decay <= 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow_params = [p.detach().clone() for p in _trainable(parameters)]
        self.collected_params = []

    def _current_decay(self):
        if self.num_updates is None:
            return self.decay
        self.num_updates += 1
        n = self.num_updates
        return min(self.decay, (1 + n) / (10 + n))

    @torch.no_grad()
    def update(self, parameters):
        decay = self._current_decay()
        params = _trainable(parameters)
        if self.shadow_params and self.shadow_params[0].is_cuda:
            from ... import train_ops
            if not hasattr(self, "_tables"):
                self._tables = train_ops._Tables()
            train_ops.ema_update(self.shadow_params, params, decay, self._tables)
            return
        delta = torch._foreach_sub(self.shadow_params, params)
        torch._foreach_mul_(delta, 1.0 - decay)
        torch._foreach_sub_(self.shadow_params, delta)

    def copy_to(self, parameters):
        """Overwrites the trainable parameters with their averages (in place, so engine-side change detection sees it)."""
        for avg, p in zip(self.shadow_params, _trainable(parameters)):
            p.data.copy_(avg.data)

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):
        for saved, p in zip(self.collected_params, parameters):
            p.data.copy_(saved.data)

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params}

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]
        loaded = state_dict["shadow_params"]
        if self.shadow_params:  # keep every average on the device of the tensor it shadows
            loaded = [s.to(old.device) if isinstance(s, torch.Tensor) else s for s, old in zip(loaded, self.shadow_params)]
        self.shadow_params = loaded
