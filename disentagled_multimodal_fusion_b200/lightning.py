"""Minimal stand-in for the pytorch_lightning / torchmetrics surface the reference relies on.

The reference modules are ``pl.LightningModule``s driven by ``pl.Trainer`` (run.py:199-247);
neither package is installable in this image, so the hot-path mirror ships the small subset of
semantics the reference actually uses: automatic optimisation, train/eval toggling,
validation after every train epoch, ``on_train_epoch_end`` (drives ``annealing_step``), per-epoch
LR schedulers (ReduceLROnPlateau fed the logged ``val_loss``), ``self.trainer.max_epochs``,
``trainer.test`` and ``save_checkpoint`` (a dict with ``state_dict``).  If the real
pytorch_lightning is importable it is used instead.
"""
from __future__ import annotations

import random
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

try:  # pragma: no cover - not available in the build image
    import pytorch_lightning as _pl
    LightningModule = _pl.LightningModule
    Trainer = _pl.Trainer
    seed_everything = _pl.seed_everything
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    def seed_everything(seed: int, workers: bool = False) -> int:
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        return seed

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.trainer = None
            self._logged: Dict[str, Any] = {}
            self.hparams: Dict[str, Any] = {}

        def save_hyperparameters(self, *a, ignore=None, **k):
            pass

        def log(self, name, value, *a, **k):
            self._logged[name] = value

        def log_dict(self, d, *a, **k):
            self._logged.update(d)

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        # hooks the reference overrides
        def on_train_epoch_end(self):
            pass

        def on_validation_epoch_end(self):
            pass

        def on_test_epoch_end(self):
            pass

    def _to_device(batch, device):
        if torch.is_tensor(batch):
            return batch.to(device, non_blocking=True)
        if isinstance(batch, (list, tuple)):
            return type(batch)(_to_device(b, device) for b in batch)
        return batch

    class Trainer:
        """fit / validate / test / save_checkpoint with Lightning's automatic-optimisation order."""

        def __init__(self, max_epochs: int = 1, accelerator: str = "auto", devices=None, logger=None,
                     enable_progress_bar: bool = False, enable_model_summary: bool = False,
                     log_every_n_steps: int = 50, **_):
            self.max_epochs = max_epochs
            self.logger = logger
            self.model: Optional[LightningModule] = None
            self.callback_metrics: Dict[str, Any] = {}
            self.last_lr: Optional[float] = None
            self.device = torch.device("cuda", torch.cuda.current_device()) if (
                accelerator in ("auto", "gpu", "cuda") and torch.cuda.is_available()) else torch.device("cpu")

        def _configure(self, model):
            cfg = model.configure_optimizers()
            sched, monitor = None, None
            if isinstance(cfg, dict):
                opt = cfg["optimizer"]
                ls = cfg.get("lr_scheduler")
                monitor = cfg.get("monitor")
                if isinstance(ls, dict):
                    sched = ls.get("scheduler")
                    monitor = ls.get("monitor", monitor)
                else:
                    sched = ls
            else:
                opt = cfg
            return opt, sched, monitor

        def fit(self, model, train_dataloaders=None, val_dataloaders=None):
            self.model = model
            model.trainer = self
            model.to(self.device)
            opt, sched, monitor = self._configure(model)
            for _epoch in range(self.max_epochs):
                model.train()
                for bi, batch in enumerate(train_dataloaders):
                    batch = _to_device(batch, self.device)
                    loss = model.training_step(batch, bi)
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    opt.step()
                if val_dataloaders is not None and hasattr(model, "validation_step"):   # Lightning skips it too
                    self._eval_loop(model, val_dataloaders, "validation")
                model.on_train_epoch_end()
                self.callback_metrics.update(model._logged)
                if sched is not None:
                    if isinstance(sched, torch.optim.lr_scheduler.ReduceLROnPlateau):
                        m = model._logged.get(monitor or "val_loss")
                        if m is not None:
                            sched.step(float(m))
                    else:
                        sched.step()
                self.last_lr = float(opt.param_groups[0]["lr"])
            return self

        @torch.no_grad()
        def _eval_loop(self, model, loader, kind):
            was_training = model.training
            model.eval()
            step = model.validation_step if kind == "validation" else model.test_step
            for bi, batch in enumerate(loader):
                step(_to_device(batch, self.device), bi)
            (model.on_validation_epoch_end if kind == "validation" else model.on_test_epoch_end)()
            self.callback_metrics.update(model._logged)
            model.train(was_training)

        def validate(self, model, dataloaders=None, verbose=False):
            model.trainer = self
            model.to(self.device)
            self._eval_loop(model, dataloaders, "validation")
            return [dict(model._logged)]

        def test(self, model, dataloaders=None, verbose=False):
            model.trainer = self
            model.to(self.device)
            self._eval_loop(model, dataloaders, "test")
            return [{k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else v)
                     for k, v in model._logged.items() if k.startswith("test")}]

        def save_checkpoint(self, path: str):
            import os
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            torch.save({"state_dict": self.model.state_dict(), "epoch": self.max_epochs}, path)


class Accuracy(nn.Module):
    """torchmetrics.Accuracy(task='multiclass') subset: forward/update/compute/reset.
    Counters stay on the device (no per-step host sync)."""

    def __init__(self, task: str = "multiclass", num_classes: Optional[int] = None, **_):
        super().__init__()
        self.num_classes = num_classes
        self.register_buffer("correct", torch.zeros((), dtype=torch.long), persistent=False)
        self.register_buffer("total", torch.zeros((), dtype=torch.long), persistent=False)

    @staticmethod
    def _preds(preds, target):
        return preds.argmax(dim=-1) if preds.ndim == target.ndim + 1 else preds

    def update(self, preds, target):
        p = self._preds(preds, target)
        self.correct += (p == target).sum()
        self.total += target.numel()

    def forward(self, preds, target):
        p = self._preds(preds, target)
        self.update(p, target)
        return (p == target).float().mean()

    def compute(self):
        return self.correct.float() / self.total.clamp(min=1).float()

    def reset(self):
        self.correct.zero_()
        self.total.zero_()
