"""Fused flat-buffer Adam / AdamW behind the ``torch.optim.Optimizer`` surface.

The reference's ``configure_optimizers`` (models/dmvae.py:204-210, models/disentangledssl.py:200-210,
models/evidential_probe.py:205-212,399-408, models/baselines.py:139-150) return ``torch.optim.Adam`` / ``AdamW`` plus a
``torch.optim.lr_scheduler``.  ``FusedAdam`` keeps that contract -- ``param_groups`` with an ``lr`` the stock schedulers
drive, ``zero_grad`` / ``step`` / ``state_dict`` -- while the update itself is ONE ``dmf_adam_step`` launch over a flat
parameter buffer (``dp.FlatParams``) and, under torch.distributed, ONE gradient all-reduce per step: a user of
``training_step`` + ``optimizer.step()`` runs the same optimizer path ``bench.py`` times.
"""
from __future__ import annotations

import torch

from . import dp


class FusedAdam(torch.optim.Optimizer):
    """Adam (``decoupled=False``: L2 ``weight_decay`` added to the gradient, torch.optim.Adam semantics) or AdamW
    (``decoupled=True``).  All parameters form one group; they must live on one CUDA device."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 allreduce: bool = True):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam: no trainable parameters")
        if any(not p.is_cuda for p in params):
            raise ValueError("FusedAdam needs CUDA parameters (there is no CPU path)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        self.flat = dp.FlatParams(params)          # re-points p.data / p.grad into the flat buffers
        self.allreduce = allreduce

    def zero_grad(self, set_to_none: bool = False):     # grads stay views of the flat buffer
        self.flat.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        if self.allreduce:
            self.flat.allreduce_grads()                 # no-op without torch.distributed
        self.flat.adam_step(float(g["lr"]), betas=tuple(g["betas"]), eps=g["eps"], weight_decay=g["weight_decay"],
                            decoupled=g["decoupled"])
        return loss

    def state_dict(self):
        return {"state": {"step": self.flat.step_count, "exp_avg": self.flat.m.clone(), "exp_avg_sq": self.flat.v.clone()},
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        st = sd["state"]
        self.flat.step_count = int(st["step"])
        self.flat.m.copy_(st["exp_avg"])
        self.flat.v.copy_(st["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd.get("param_groups", [])):
            g.update({k: v for k, v in s.items() if k != "params"})


class FusedAdamW(FusedAdam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=True, **kw)


def make_optimizer(cls, params, **kw):
    """``cls(params, **kw)`` with torch.optim.Adam / AdamW replaced by their fused flat-buffer equivalents when the
    parameters are on a CUDA device (any other optimizer class, or CPU parameters, is passed through unchanged)."""
    params = list(params)
    on_gpu = bool(params) and all(p.is_cuda for p in params if p.requires_grad)
    if on_gpu and cls is torch.optim.Adam:
        return FusedAdam(params, **kw)
    if on_gpu and cls is torch.optim.AdamW:
        kw.setdefault("weight_decay", 1e-2)
        return FusedAdamW(params, **kw)
    return cls(params, **kw)
