"""Host-side helpers of the reference's ``run.py`` that the drop-in surface keeps: the dot-path
config getter ``C`` (run.py:29-36), ``_get_dataset`` (:38-50), ``_split_indices`` (:52-56), the data helpers
``get_normal_data`` / ``get_conflict_data`` (:59-102) and ``build_factories`` (:135-175).  The loaders they return
are device-resident (``datasets.DeviceLoader``) when a GPU is present and ``kernels.device_resident_data`` is set,
else plain ``torch.utils.data.DataLoader`` objects exactly like the reference."""
from __future__ import annotations

import os
from functools import partial
from pathlib import Path

import numpy as np
import torch
import yaml
from torch.utils.data import DataLoader, Subset

from . import baselines
from .datasets import CUB, PIE, Caltech, DeviceLoader, HandWritten, Scene
from .dmvae import DMVAE
from .evidential_probe import DisentangledEvidentialProbeModule, EvidentialProbeModule

CFG_PATH = Path(os.environ.get("DMF_CONFIG", Path(__file__).resolve().parent.parent / "configs" / "config.yaml"))
cfg = yaml.safe_load(open(CFG_PATH)) if CFG_PATH.is_file() else {}


def load_config(path):
    """(Re)load the YAML the dot-path getter ``C`` reads -- e.g. one of the reference's own ``configs/*.yaml``, which
    parse unchanged (including the pasted-text quirk of synthetic_config.yaml that turns its ``data`` block into a
    top-level ``yamldata`` key, so that every ``C('data.common_med.*')`` falls back to its in-code default)."""
    global cfg, CFG_PATH
    CFG_PATH = Path(path)
    cfg = yaml.safe_load(open(CFG_PATH)) or {}
    return cfg


def C(path, default=None):
    """Dot-path getter with default: ``C('probes.dropout_p', 0.1)``."""
    cur = cfg
    for key in path.split("."):
        if not isinstance(cur, dict) or key not in cur:
            return default
        cur = cur[key]
    return cur


_DATASETS = {"CUB": CUB, "CalTech": Caltech, "HandWritten": HandWritten, "PIE": PIE, "Scene": Scene}


def _get_dataset(dataset_name):
    if dataset_name not in _DATASETS:
        raise ValueError(f"Unknown dataset: {dataset_name}")
    return _DATASETS[dataset_name]()


def _split_indices(n, train_frac):
    idx = np.arange(n)
    np.random.shuffle(idx)            # numpy RNG seeded by the caller (pl.seed_everything in the reference)
    n_train = int(train_frac * n)
    return idx[:n_train], idx[n_train:]


def _loaders(dataset, train_idx, test_idx):
    bs = C("dataloader.batch_size", 100)
    if C("kernels.device_resident_data", True) and torch.cuda.is_available():
        return (DeviceLoader.from_dataset(dataset, bs, indices=train_idx, shuffle=True),
                DeviceLoader.from_dataset(dataset, bs, indices=test_idx, shuffle=False))
    nw = C("dataloader.num_workers", 0)
    return (DataLoader(Subset(dataset, train_idx), batch_size=bs, shuffle=True, num_workers=nw),
            DataLoader(Subset(dataset, test_idx), batch_size=bs, shuffle=False, num_workers=nw))


def get_normal_data(dataset_name):
    dataset = _get_dataset(dataset_name)
    train_idx, test_idx = _split_indices(len(dataset), C("data.split.train_frac", 0.8))
    tl, vl = _loaders(dataset, train_idx, test_idx)
    return tl, vl, dataset.num_classes, dataset.num_views, list(np.squeeze(dataset.dims))


def get_conflict_data(dataset_name):
    dataset = _get_dataset(dataset_name)
    train_idx, test_idx = _split_indices(len(dataset), C("data.split.train_frac", 0.8))
    pp = C("data.conflict", {})
    dataset.postprocessing(test_idx, addNoise=pp.get("addNoise", False), sigma=pp.get("sigma", 0.5),
                           ratio_noise=pp.get("ratio_noise", 0.0), addConflict=pp.get("addConflict", True),
                           ratio_conflict=pp.get("ratio_conflict", 1.0))
    tl, vl = _loaders(dataset, train_idx, test_idx)      # built AFTER the corruption of the test rows
    return tl, vl, dataset.num_classes, dataset.num_views, list(np.squeeze(dataset.dims))


def build_factories(model_params, probe_input_dim, dmvae_kwargs):
    """(DMVAEFactory, ProbeFactory, DisProbeFactory, LateFusionFactory) with the reference's argument wiring."""
    common = dict(num_classes=model_params["classes"], lr=model_params["lr"],
                  annealing_start=model_params["annealing_start"], hidden_dim=model_params["model_hidden_dim"],
                  dropout=model_params["dropout_p"])
    DMVAEFactory = partial(DMVAE, feature_encoders=model_params["classifiers"], output_dim=model_params["output_dims"],
                           dropout=dmvae_kwargs["dropout"], a=dmvae_kwargs["a"], hidden_dim=dmvae_kwargs["hidden_dim"],
                           embed_dim=dmvae_kwargs["embed_dim"], lr=dmvae_kwargs["lr"], num_epochs=dmvae_kwargs["num_epochs"])
    ProbeFactory = partial(EvidentialProbeModule, input_dim=probe_input_dim, **common)
    DisProbeFactory = partial(DisentangledEvidentialProbeModule, input_dim=probe_input_dim, **common)
    LateFusionFactory = partial(baselines.LateFusion, model_params["classifiers"], model_params["output_dims"],
                                model_params["classes"], dropout=model_params["dropout_p"], lr=model_params["lr"],
                                annealing_start=model_params["annealing_start"], hidden_dim=model_params["model_hidden_dim"])
    return DMVAEFactory, ProbeFactory, DisProbeFactory, LateFusionFactory
