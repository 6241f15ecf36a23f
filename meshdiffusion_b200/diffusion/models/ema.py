"""Exponential moving average of the trainable parameters (reference: lib/diffusion/models/ema.py:10-98).

Same arithmetic (`decay_t = min(decay, (1+n)/(10+n))`, `s -= (1-decay_t)(s-p)`) and the same positional
`state_dict` layout {'decay','num_updates','shadow_params'} so reference checkpoints interchange.
"""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow_params = [p.clone().detach() for p in parameters if p.requires_grad]
        self.collected_params = []

    def update(self, parameters):
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        with torch.no_grad():
            params = [p for p in parameters if p.requires_grad]
            # one multi-tensor pass instead of 494 small kernels
            diffs = torch._foreach_sub(self.shadow_params, params)
            torch._foreach_mul_(diffs, one_minus_decay)
            torch._foreach_sub_(self.shadow_params, diffs)

    def copy_to(self, parameters):
        params = [p for p in parameters if p.requires_grad]
        for s, p in zip(self.shadow_params, params):
            p.data.copy_(s.data)

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):
        for c, p in zip(self.collected_params, parameters):
            p.data.copy_(c.data)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]
        self.shadow_params = [s.to(p.device) if isinstance(s, torch.Tensor) else s
                              for s, p in zip(state_dict["shadow_params"], self.shadow_params)] \
            if self.shadow_params else state_dict["shadow_params"]
