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


def assert_close(a, b, rtol, what="", atol=0.0):
    """max-norm relative error (|a-b|_inf / |b|_inf) -- the tolerance BASELINE.json states is
    relative to the magnitude of the quantity (losses, embeddings, gradients).  ``atol`` is only
    used where the quantity itself is a cancellation residue (e.g. an InfoNCE loss of 2e-6 formed
    from logits of magnitude 1/T = 14.3, where fp32 rounding of the logits alone is 1.7e-6)."""
    a64 = torch.as_tensor(a).detach().to(torch.float64).cpu()
    b64 = torch.as_tensor(b).detach().to(torch.float64).cpu()
    err = float((a64 - b64).abs().max())
    ref = float(b64.abs().max())
    assert err <= atol + rtol * ref, f"{what}: abs err {err:.3e} (rel {err / (ref + 1e-30):.3e}) > {atol:.1e} + {rtol:.1e}*{ref:.3e}"


def mlp_params_from_sd(sd, prefix, idxs, device="cpu", grad=False):
    ws = [T(sd[f"{prefix}.layers.{i}.weight"], device, grad) for i in idxs]
    bs = [T(sd[f"{prefix}.layers.{i}.bias"], device, grad) for i in idxs]
    return ws, bs


def check_sampled_grads(named_grads, g, tol):
    """fixtures at the BASELINE sizes keep, per parameter, the gradient norm and 64 sampled entries"""
    for k, gr in named_grads.items():
        flat = gr.reshape(-1)
        ref_norm = float(g["gnorm." + k])
        assert abs(float(flat.double().norm()) - ref_norm) <= tol * max(ref_norm, 1e-12), ("gnorm", k)
        idx = torch.from_numpy(g["gidx." + k]).to(flat.device)
        ref = T(g["gval." + k], flat.device)
        scale = max(float(g["gmax." + k]), 1e-20)      # max-norm relative error, like assert_close on a full tensor
        err = float((flat[idx] - ref).abs().max())
        assert err <= tol * scale, ("gval", k, err, scale)
