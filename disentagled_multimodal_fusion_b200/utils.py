"""Host-side mirror of the reference's ``utils.py`` (same names, arguments and semantics); the
tensor math runs in the fused K3 kernel (``ops.fuse_evidence`` / ``dmf_edl_fused``)."""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import lib, check, ptr, stream


# ---- schedulers (utils.py:10-42) -- host scalars
class Scheduler:
    def __call__(self, **kwargs):
        raise NotImplementedError()


class LinearScheduler(Scheduler):
    def __init__(self, start_value, end_value, n_iterations, start_iteration=0):
        self.start_value, self.end_value = start_value, end_value
        self.n_iterations, self.start_iteration = n_iterations, start_iteration
        self.m = (end_value - start_value) / n_iterations

    def __call__(self, iteration):
        if iteration > self.start_iteration + self.n_iterations:
            return self.end_value
        if iteration <= self.start_iteration:
            return self.start_value
        return (iteration - self.start_iteration) * self.m + self.start_value


class ExponentialScheduler(LinearScheduler):
    def __init__(self, start_value, end_value, n_iterations, start_iteration=0, base=10):
        self.base = base
        super().__init__(math.log(start_value, base), math.log(end_value, base), n_iterations, start_iteration)

    def __call__(self, iteration):
        return self.base ** super().__call__(iteration)


# ---- evidence activation (utils.py:46-63)
class _Evidence(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        h = h.float().contiguous()
        e = torch.empty_like(h)
        check(lib.dmf_evidence_fwd(ptr(h), ptr(e), h.numel(), stream()))
        ctx.save_for_backward(h, e)
        return e

    @staticmethod
    def backward(ctx, de):
        h, e = ctx.saved_tensors
        de = de.float().contiguous()
        dh = torch.empty_like(h)
        check(lib.dmf_evidence_bwd(ptr(h), ptr(e), ptr(de), ptr(dh), h.numel(), stream()))
        return dh


def activation_function(h, activation="exp"):
    if activation == "softplus":
        return nn.functional.softplus(h)
    from . import _lib
    _lib.require_device()
    return _Evidence.apply(h)


# ---- fusion rules (utils.py:66-116): all five run inside the K3 kernel
def get_cml_fusion(all_evidences):
    return ops.fuse_evidence(all_evidences, "cml")


def get_avg_fusion(all_evidences):
    return ops.fuse_evidence(all_evidences, "avg")


def get_disentangled_fusion(all_evidences, shared_index=0):
    if shared_index != 0:
        raise ValueError("the fused kernel implements the reference default shared_index=0")
    return ops.fuse_evidence(all_evidences, "disentangled")


def get_joint_fusion(all_evidences, shared_index=0, shared_weight=0.5):
    if shared_index != 0 or shared_weight != 0.5:
        raise ValueError("the fused kernel implements the reference defaults shared_index=0, shared_weight=0.5")
    return ops.fuse_evidence(all_evidences, "joint")


def discounted_belief_fusion(all_evidences, flambda=3):
    if flambda != 3:
        raise ValueError("the fused kernel implements the reference default flambda=3")
    return ops.fuse_evidence(all_evidences, "dbf")


FUSION_NAME = {get_cml_fusion: "cml", get_avg_fusion: "avg", get_disentangled_fusion: "disentangled",
               get_joint_fusion: "joint", discounted_belief_fusion: "dbf"}


# ---- init (utils.py:153-166)
def initialize_weights(model, initialization="xavier"):
    for m in model.modules():
        if isinstance(m, nn.Linear):
            if initialization == "xavier":
                nn.init.xavier_uniform_(m.weight)
            elif initialization == "zeros":
                nn.init.zeros_(m.weight)
            elif initialization == "normal":
                nn.init.normal_(m.weight, mean=0, std=0.01)
            elif initialization == "uniform":
                nn.init.uniform_(m.weight, a=-0.05, b=0.05)
            else:
                raise NotImplementedError()
    return model


# ---- augmentation (utils.py:118-151): runs BEFORE forward; kept on the host generator like the
# reference (numpy RNG for the per-row choice), vectorised instead of a per-row Python loop.
def augment_data(x_batch, noise_scale=0.01, drop_scale=10):
    B, D = x_batch.shape
    v2 = torch.clone(x_batch)
    for i in range(B):        # same per-row draw order as the reference (numpy choice, then torch/numpy noise)
        t = np.random.choice(3, 1, replace=False)[0]
        if t == 0:
            v2[i] = v2[i] + (torch.randn(v2[i].shape) * noise_scale).to(v2.device)
        elif t == 1:
            idx = np.random.choice(D, D // drop_scale, replace=False)
            v2[i, idx] = 0.0
    return v2


# ---- vMF noise on the host generator, replaying the reference's draw order
def draw_vmf_noise(B: int, D: int, kappa: float = 1.0, k: int = 1, dtype=torch.float32):
    """(w [B,1], v [B,D-1]) drawn exactly like VonMisesFisher.rsample does on the CPU generator
    (models/classifiers.py:314-431): rejection loop {Beta(fp64), Uniform}, then Normal [B,D] with the
    first column dropped and rows normalised.  Same seed => same noise as the reference."""
    m = D
    scale = kappa * torch.ones(B, 1, dtype=dtype)
    c = torch.sqrt((4 * (scale ** 2)) + (m - 1) ** 2)
    b_true = (-2 * scale + c) / (m - 1)
    b_app = (m - 1) / (4 * scale)
    s = torch.min(torch.max(torch.tensor([0.0], dtype=dtype), scale - 10), torch.tensor([1.0], dtype=dtype))
    b = b_app * s + b_true * (1 - s)
    a = (m - 1 + 2 * scale + c) / 4
    d = (4 * a * b) / (1 + b) - (m - 1) * math.log(m - 1)
    w = torch.zeros_like(b)
    active = torch.ones_like(b) == 1
    eps = 1e-20
    while active.sum() != 0:
        con = torch.tensor((m - 1) / 2, dtype=torch.float64)
        e_ = torch.distributions.Beta(con, con).sample(torch.Size([B, k])).type(dtype)
        u = torch.distributions.Uniform(0 + eps, 1 - eps).sample(torch.Size([B, k])).type(dtype)
        w_ = (1 - (1 + b) * e_) / (1 - (1 - b) * e_)
        t = (2 * a * b) / (1 - (1 - b) * e_)
        accept = ((m - 1.0) * t.log() - t + d) > torch.log(u)
        any_acc = accept.any(dim=1, keepdim=True)
        idx = accept.float().argmax(dim=1, keepdim=True)
        take = active & any_acc
        w[take] = w_.gather(1, idx)[take]
        active = active & ~any_acc
    v = torch.distributions.Normal(0, 1).sample(torch.Size([B, D])).type(dtype)[:, 1:]
    v = v / v.norm(dim=-1, keepdim=True)
    return w, v
