"""Size-independent properties of the CPU oracle (oracle/port.py) -- the checker the GPU parity tests lean on.
The golden fixtures pin it to the unmodified reference at fixed inputs (tests/test_oracle_golden.py); these
properties pin its STRUCTURE on random inputs: closed forms, symmetries and limits that the reference's formulas
(utils.py:46-116, models/losses.py:17-248, models/classifiers.py:314-437) imply.  float64 so the assertions are sharp."""
import math

import pytest

import torch
from hypothesis import given, settings, strategies as st

from oracle import port

DT = torch.float64
SET = dict(max_examples=25, deadline=None, derandomize=True)      # fixed example sequence: no run-to-run flakiness


def _evid(seed, B, V, C, scale=1.5):
    g = torch.Generator().manual_seed(seed)
    return torch.exp(torch.randn(B, V, C, generator=g, dtype=DT) * scale), torch.randint(0, C, (B,), generator=g)


@settings(**SET)
@given(seed=st.integers(0, 10**6), B=st.integers(1, 9), V=st.integers(1, 6), C=st.integers(2, 12))
def test_fusion_rules_closed_forms_and_symmetries(seed, B, V, C):
    e, _ = _evid(seed, B, V, C)
    cml, avg = port.fuse(e, "cml"), port.fuse(e, "avg")
    assert torch.allclose(avg * V, cml)
    if V > 1:
        assert torch.allclose(port.fuse(e, "joint"), 0.5 * e[:, 0] + 0.5 * port.fuse(e, "disentangled"))
        perm = torch.cat([torch.zeros(1, dtype=torch.long), 1 + torch.randperm(V - 1)])      # view 0 is the shared head
        for agg in ("cml", "avg", "joint", "disentangled"):
            assert torch.allclose(port.fuse(e[:, perm], agg), port.fuse(e, agg))
    # class permutation commutes with every rule (dbf included: it only sums / compares over the class axis)
    cp = torch.randperm(C)
    for agg in ("cml", "avg", "dbf"):
        assert torch.allclose(port.fuse(e[:, :, cp], agg), port.fuse(e, agg)[:, cp], rtol=1e-9, atol=1e-12)
    # dbf with identical views: no conflict -> no discount -> C b / (u + 1e-6) of the common opinion
    same = e[:, :1].expand(B, V, C).contiguous()
    S = (same[:, 0] + 1).sum(-1, keepdim=True)
    assert torch.allclose(port.fuse(same, "dbf"), C * (same[:, 0] / S) / (C / S + 1e-6), rtol=1e-9)


@settings(**SET)
@given(seed=st.integers(0, 10**6), B=st.integers(1, 8), V=st.integers(1, 5), C=st.integers(2, 10),
       step=st.integers(0, 30), fused=st.sampled_from([0.0, 1.0]))
def test_avg_trusted_loss_structure(seed, B, V, C, step, fused):
    e, y = _evid(seed, B, V, C)
    loss = port.avg_trusted_loss(e, y, None, fused, step, 10)
    # view permutation and joint (class, label) permutation leave the loss unchanged
    assert torch.allclose(port.avg_trusted_loss(e[:, torch.randperm(V)], y, None, fused, step, 10), loss, rtol=1e-10)
    cp = torch.randperm(C)
    inv = torch.empty_like(cp)
    inv[cp] = torch.arange(C)
    assert torch.allclose(port.avg_trusted_loss(e[:, :, cp], inv[y], None, fused, step, 10), loss, rtol=1e-10)
    # the aggregated evidence is not used and the conflict term is linear in `fused`
    assert torch.equal(port.avg_trusted_loss(e, y, port.fuse(e, "avg"), fused, step, 10), loss)
    t = min(1.0, step / 10)
    base = port.avg_trusted_loss(e, y, None, 0.0, step, 10)
    assert torch.allclose(loss, base + fused * (0.2 * (1 - t) + t) * port.dc_loss(e), rtol=1e-10)
    # annealing: coef saturates at 1 once step >= annealing_start; at step 0 only the digamma term is left
    if step >= 10:
        assert torch.allclose(base, port.avg_trusted_loss(e, y, None, 0.0, 10, 10))
    alpha = (e + 1).reshape(B * V, C)
    yy = y.repeat_interleave(V)
    A = (torch.digamma(alpha.sum(1)) - torch.digamma(alpha[torch.arange(B * V), yy])).mean() / V
    assert torch.allclose(port.avg_trusted_loss(e, y, None, 0.0, 0, 10), A, rtol=1e-10)
    # closed form of the backward that the fused kernel implements
    ev = e.clone().requires_grad_()
    (g,) = torch.autograd.grad(port.avg_trusted_loss(ev, y, None, 0.0, step, 10), ev)
    coef = float(torch.tensor(min(1.0, step / 10), dtype=torch.float32))     # the reference forms coef in fp32 (losses.py:127-130)
    al = e + 1
    Ssum = al.sum(-1, keepdim=True)
    y1 = torch.nn.functional.one_hot(y, C).to(DT)[:, None, :].expand(B, V, C)
    ay = (al * y1).sum(-1, keepdim=True)
    St = Ssum - ay + 1
    tri = lambda x: torch.polygamma(1, x)                                       # noqa: E731
    g_other = tri(Ssum) + coef * ((al - 1) * tri(al) - (St - C) * tri(St))
    g_label = tri(Ssum) - tri(ay)
    ref = (y1 * g_label + (1 - y1) * g_other) / (B * V * V)
    assert torch.allclose(g, ref, rtol=1e-8, atol=1e-14)


@settings(**SET)
@given(seed=st.integers(0, 10**6), B=st.integers(1, 8), V=st.integers(2, 5), C=st.integers(2, 10))
def test_dc_loss_bounds_and_zero(seed, B, V, C):
    e, _ = _evid(seed, B, V, C)
    dc = port.dc_loss(e)
    assert 0.0 <= float(dc) <= V + 1e-12          # each pair term is a total-variation distance times (1-u)(1-u') <= 1
    same = e[:, :1].expand(B, V, C).contiguous()
    assert float(port.dc_loss(same)) < 1e-14


@settings(**SET)
@given(seed=st.integers(0, 10**6), B=st.integers(2, 12), D=st.integers(2, 16))
def test_supcon_symmetries_and_clip_form(seed, B, D):
    g = torch.Generator().manual_seed(seed)
    z0 = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=DT), dim=-1)
    # positives kept close to their anchors: the reference adds 1e-12 inside the log after subtracting the row max
    # (the self-similarity 1/T = 14.3), so rows whose cross-view logits are ALL below -27 deviate from the closed forms
    z1 = torch.nn.functional.normalize(z0 + 0.15 * torch.randn(B, D, generator=g, dtype=DT), dim=-1)
    loss, lx, ly = port.supcon(z0, z1)
    l2, lx2, ly2 = port.supcon(z1, z0)                       # swapping the views swaps the diagnostics only
    assert torch.allclose(l2, loss) and torch.allclose(lx2, ly) and torch.allclose(ly2, lx)
    p = torch.randperm(B)                                     # batch order is irrelevant
    for a, b in zip(port.supcon(z0[p], z1[p]), (loss, lx, ly)):
        assert torch.allclose(a, b, rtol=1e-10)
    q, _ = torch.linalg.qr(torch.randn(D, D, generator=g, dtype=DT))   # common rotation
    for a, b in zip(port.supcon(z0 @ q, z1 @ q), (loss, lx, ly)):
        assert torch.allclose(a, b, rtol=1e-9)
    # symmetric CLIP InfoNCE: 0.5 (CE(S01, diag) + CE(S01^T, diag)) with S01 = z0 z1^T / T   
    S = z0 @ z1.T / 0.07
    tgt = torch.arange(B)
    clip = 0.5 * (torch.nn.functional.cross_entropy(S, tgt) + torch.nn.functional.cross_entropy(S.T, tgt))
    assert torch.allclose(loss, clip, rtol=1e-6)
    # gradient closed form the tiled backward implements: dz0 = [(P01 - I) z1 + (P10 - I)^T z1] / (2 B T)
    a = z0.clone().requires_grad_()
    (ga,) = torch.autograd.grad(port.supcon(a, z1)[0], a)
    P01, P10 = torch.softmax(S, 1), torch.softmax(S.T, 1)
    eye = torch.eye(B, dtype=DT)
    assert torch.allclose(ga, ((P01 - eye) @ z1 + (P10 - eye).T @ z1) / (2 * B * 0.07), rtol=1e-6, atol=1e-8)


@settings(**SET)
@given(seed=st.integers(0, 10**6), B=st.integers(1, 8), D=st.integers(3, 24))
def test_vmf_sample_is_a_unit_vector_and_reflection(seed, B, D):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(B, D, generator=g, dtype=DT)
    w = torch.rand(B, 1, generator=g, dtype=DT) * 2 - 1
    v = torch.nn.functional.normalize(torch.randn(B, D - 1, generator=g, dtype=DT), dim=-1)
    z = port.vmf_rsample(e, w, v)
    # the reflection vector is normalised with ||.|| + 1e-5, so the Householder map is orthogonal only up to
    # ~4e-5 / ||e1 - loc||
    n = (torch.nn.functional.one_hot(torch.zeros(B, dtype=torch.long), D).to(DT) - e / e.norm(dim=-1, keepdim=True)).norm(dim=-1)
    assert torch.all((z.norm(dim=-1) - 1).abs() <= 1e-4 / n + 1e-9)
    # w = 1 puts the sample on the mode: z = loc (up to the same 1e-5 regulariser)
    z1 = port.vmf_rsample(e, torch.ones(B, 1, dtype=DT), v)
    assert torch.all((z1 - e / e.norm(dim=-1, keepdim=True)).norm(dim=-1) <= 1e-4 / n + 1e-6)


def test_evidence_activation_limits():
    h = torch.linspace(-20, 20, 401, dtype=DT)
    ev = port.evidence_activation(h)
    assert torch.all(ev[1:] >= ev[:-1])                                      # monotone
    assert torch.allclose(ev[h <= -10], torch.full_like(ev[h <= -10], math.exp(-10)), rtol=1e-9)
    assert torch.allclose(ev[h >= 10], ev[h == 10].expand_as(ev[h >= 10]))   # clamped above
    mid = (h > -9) & (h < 9)
    assert torch.allclose(ev[mid], torch.exp(h[mid]), rtol=1e-8)             # the soft cap at 1e13 is far away


def test_uncertainty_summaries_limits():
    C = 7
    vac = torch.zeros(3, C, dtype=DT)                                        # no evidence: total uncertainty
    u, ale, arg = port.uncertainty_summaries(vac)
    assert torch.allclose(u, torch.ones(3, dtype=DT))
    conf = torch.zeros(1, C, dtype=DT)
    conf[0, 4] = 1e6
    u2, ale2, arg2 = port.uncertainty_summaries(conf)
    assert float(u2) < 1e-4 and int(arg2) == 4 and float(ale2) < float(ale[0])


def test_stored_probability_form_of_the_supcon_backward():
    """e (fa + fb) reproduces the autograd gradient of the reference formula for unit-norm views (the identity the
    stored-probability kernels rely on), and the blocked E layout round-trips with zero padding."""
    gen = torch.Generator().manual_seed(21)
    B, D = 70, 24
    z0 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen, dtype=torch.float64), dim=-1)
    z1 = torch.nn.functional.normalize(0.5 * z0 + 0.5 * torch.randn(B, D, generator=gen, dtype=torch.float64), dim=-1)
    a, b = z0.clone().requires_grad_(), z1.clone().requires_grad_()
    loss, _, _ = port.supcon(a, b)
    ga, gb = torch.autograd.grad(loss, (a, b))
    e, fa, fb, dz0, dz1 = port.supcon_stored_probabilities(z0, z1)
    assert float((dz0 - ga).abs().max()) < 1e-9 * float(ga.abs().max()) + 1e-15
    assert float((dz1 - gb).abs().max()) < 1e-9 * float(gb.abs().max()) + 1e-15
    assert float(e.max()) <= 1.0 + 1e-12          # |s| <= 1/T: the fixed shift is the row maximum bound
    blocks = port.infonce_e_blocks(e)
    assert blocks.shape == (2 * 4, 128, 64)
    assert torch.equal(port.infonce_e_unblock(blocks, B, B), e)
    assert float(blocks.sum()) == pytest.approx(float(e.sum()), rel=1e-12)      # padding is zero
    assert torch.equal(blocks[0, :B, :64], e[:, :64]) and float(blocks[1, :, 6:].abs().max()) == 0.0
