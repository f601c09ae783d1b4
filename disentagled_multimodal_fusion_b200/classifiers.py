"""Mirror of the hot-path classes of the reference's ``models/classifiers.py``: ``Linear`` (alias
``MLP``), ``EvidentialNN``, ``IdentityEncoder``.  Parameters live in the same ``layers`` ModuleList
(same construction / RNG order => same init for the same seed, same state_dict keys); the math runs
in the grouped-GEMM kernels (ops.grouped_mlp)."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .utils import initialize_weights


class IdentityEncoder(nn.Module):
    def forward(self, x):
        return x


def _build_layers(layers, output_dims, dropout, initialization):
    mods = nn.ModuleList()
    for i in range(len(layers) - 1):
        mods.append(nn.Linear(layers[i], layers[i + 1]))
        mods.append(nn.ReLU())
        if dropout > 0:
            mods.append(nn.Dropout(dropout))
    mods.append(nn.Linear(layers[-1], output_dims))
    return initialize_weights(mods, initialization)


class Linear(nn.Module):
    """models/classifiers.py:16-48."""
    final = "none"

    def __init__(self, dropout=0.1, output_dims=128, index=0, layers=(5, 10, 50), initialization="xavier"):
        super().__init__()
        self.dropout = dropout
        self.output_dims = output_dims
        self.layers = _build_layers(layers, output_dims, dropout, initialization)
        self.precision = "fp32"

    def linears(self):
        return [m for m in self.layers if isinstance(m, nn.Linear)]

    def weights(self):
        return [m.weight for m in self.linears()]

    def biases(self):
        return [m.bias for m in self.linears()]

    def dropout_masks(self, x):
        """Inverted-dropout masks for the hidden layers in train mode (None in eval / p = 0)."""
        if not self.training or self.dropout <= 0:
            return None
        keep = 1.0 - self.dropout
        lin = self.linears()
        return [torch.bernoulli(torch.full((x.shape[0], m.out_features), keep, device=x.device)) / keep
                for m in lin[:-1]]

    def forward(self, x):
        masks = self.dropout_masks(x)
        return ops.grouped_mlp([x], [self.weights()], [self.biases()], final=self.final, precision=self.precision,
                               dropout_masks=[masks] if masks is not None else None)[0]


MLP = Linear


class EvidentialNN(Linear):
    """models/classifiers.py:469-502: same MLP + activation_function(out, 'exp') fused in the epilogue."""
    final = "evidence"

    def __init__(self, dropout=0.1, output_dims=10, layers=(100, 100), initialization="xavier"):
        super().__init__(dropout=dropout, output_dims=output_dims, layers=layers, initialization=initialization)


def grouped_forward(mods, xs, extras=None, precision=None, opts=None):
    """Run several same-depth MLP modules in one launch per layer (views x streams)."""
    masks = [m.dropout_masks(x) for m, x in zip(mods, xs)]
    if all(mk is None for mk in masks):
        masks = None
    return ops.grouped_mlp(list(xs), [m.weights() for m in mods], [m.biases() for m in mods], final=mods[0].final,
                           precision=precision or mods[0].precision, dropout_masks=masks, extras=extras, opts=opts)
