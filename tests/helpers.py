"""Shared helpers for the parity tests (golden loading, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def T(a, device="cpu", grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if grad:
        t.requires_grad_()
    return t


def relerr(a, b):
    a = torch.as_tensor(a).detach().to(torch.float64).cpu()
    b = torch.as_tensor(b).detach().to(torch.float64).cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, rtol, what=""):
    """max-norm relative error (|a-b|_inf / |b|_inf) -- the tolerance BASELINE.json states is
    relative to the magnitude of the quantity (losses, embeddings, gradients)."""
    e = relerr(a, b)
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:.1e}"


def mlp_params_from_sd(sd, prefix, idxs, device="cpu", grad=False):
    ws = [T(sd[f"{prefix}.layers.{i}.weight"], device, grad) for i in idxs]
    bs = [T(sd[f"{prefix}.layers.{i}.bias"], device, grad) for i in idxs]
    return ws, bs
