"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

This file is part of ``oracle/``: only ``tests/``, ``tests/golden/make_golden.py``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline/reference legs may import
it.  It never ships on the product path.

The reference (``/root/reference``) is a pure-Python PyTorch-Lightning repo.  It can be
imported in the build container (never on the GPU box, where the tree is absent) once
two third-party packages that are not installed here are stubbed:

* ``pytorch_lightning`` -> ``LightningModule`` = ``nn.Module`` with no-op logging
* ``torchmetrics``      -> minimal ``Accuracy`` / ``MeanMetric``

and three process-wide quirks are neutralised:

* ``models/dmvae.py:9`` / ``models/disentangledssl.py:13`` set
  ``torch.set_float32_matmul_precision`` to 'medium' / 'high' at import time; a 1e-5
  oracle has to undo that (``'highest'``) after importing.
* ``models/classifiers.py:461,465``, ``models/disentangledssl.py:177-178`` and
  ``utils.py:120`` call ``.cuda()`` unconditionally -> identity shim on CPU.
* the HuggingFace ``datasets`` package shadows the reference's ``datasets/`` directory
  -> load ``datasets/dataset.py`` by file path.

Nothing here copies reference source; it only executes it where it lies.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DMF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "losses.py"))


class _StubLightningModule(nn.Module):
    """Stand-in for pl.LightningModule: only what the reference modules touch."""

    def __init__(self, *a, **k):
        super().__init__()
        self.trainer = types.SimpleNamespace(max_epochs=1)

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class _StubAccuracy(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.correct = 0
        self.total = 0

    def update(self, preds, target):
        if preds.ndim == target.ndim + 1:
            preds = preds.argmax(dim=-1)
        self.correct += int((preds == target).sum())
        self.total += int(target.numel())

    def forward(self, preds, target):
        if preds.ndim == target.ndim + 1:
            preds = preds.argmax(dim=-1)
        self.update(preds, target)
        return (preds == target).float().mean()

    def compute(self):
        return torch.tensor(self.correct / max(1, self.total))

    def reset(self):
        self.correct = 0
        self.total = 0


class _StubMeanMetric(_StubAccuracy):
    pass


_loaded = None


def load_reference():
    """Import the reference modules unmodified and return them in a namespace.

    Returns a SimpleNamespace with attributes: losses, utils, classifiers, dmvae,
    disentangledssl, evidential_probe, baselines, dataset (the latter loaded by path).
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _StubLightningModule
        pl.seed_everything = lambda seed, **k: _seed_everything(seed)
        pl.Trainer = object
        loggers = types.ModuleType("pytorch_lightning.loggers")
        loggers.CSVLogger = object
        pl.loggers = loggers
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.loggers"] = loggers
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm.Accuracy = _StubAccuracy
        tm.MeanMetric = _StubMeanMetric
        sys.modules["torchmetrics"] = tm

    prev_precision = torch.get_float32_matmul_precision()
    sys.path.insert(0, REFERENCE_ROOT)
    # make sure 'utils' / 'models' resolve to the reference, not to something cached
    for name in ("utils", "models"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    try:
        ns = types.SimpleNamespace()
        ns.utils = importlib.import_module("utils")
        ns.losses = importlib.import_module("models.losses")
        ns.classifiers = importlib.import_module("models.classifiers")
        ns.dmvae = importlib.import_module("models.dmvae")
        ns.disentangledssl = importlib.import_module("models.disentangledssl")
        ns.evidential_probe = importlib.import_module("models.evidential_probe")
        ns.baselines = importlib.import_module("models.baselines")
        spec = importlib.util.spec_from_file_location(
            "dmf_reference_dataset", os.path.join(REFERENCE_ROOT, "datasets", "dataset.py"))
        ds = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ds)
        ns.dataset = ds
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # D8: undo the import-time precision change; the oracle is true fp32
        torch.set_float32_matmul_precision("highest")
        del prev_precision
    _loaded = ns
    return ns


def _seed_everything(seed: int):
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


class cuda_identity_shim:
    """Context manager: make ``Tensor.cuda()`` / ``Module.cuda()`` the identity on a CPU box
     so DisentangledSSL's hard ``.cuda()`` calls run on the host."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False


class in_reference_cwd:
    """Loaders use relative ``data/*.mat`` paths (datasets/dataset.py:275,284,295,315)."""

    def __enter__(self):
        self._cwd = os.getcwd()
        os.chdir(REFERENCE_ROOT)
        return self

    def __exit__(self, *exc):
        os.chdir(self._cwd)
        return False
