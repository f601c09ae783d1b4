"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference hot path.

Status: **pinned against outputs of the reference itself** -- the reference has no tests or
golden vectors of its own, so ``tests/golden/make_golden.py`` imports the
unmodified reference in the build container (``oracle/ref_harness.py``), runs it on seeded
inputs and commits inputs/weights/noise/outputs/gradients as fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker / the timed CPU
baseline.  The product path (``disentagled_multimodal_fusion_b200``) never imports it and has
no CPU fallback.

The arithmetic of the reference lives in PyTorch (pinned torch==2.6.0 in the reference's
requirements.txt; torch 2.11 here), so this restatement is written with torch CPU tensors and
follows the reference's op order; every function cites the reference file:line it restates.
All randomness is taken as explicit *noise arguments* (the reference draws it inline); the
``draw_*`` helpers replay the reference's draw order on the CPU generator so equal seeds give
equal noise.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# a1 / a12  MLP ("Linear") and evidential head             models/classifiers.py:16-48,469-502
# --------------------------------------------------------------------------------------
def mlp(x: Tensor, weights: Sequence[Tensor], biases: Sequence[Tensor]) -> Tensor:
    """``Linear.forward`` (models/classifiers.py:43-48): x.float(); (Linear->ReLU)* ; Linear.
    Dropout layers are identity (backbones use p=0, probes are checked in eval())."""
    h = x.to(weights[0].dtype)
    n = len(weights)
    for i in range(n):
        h = F.linear(h, weights[i], biases[i])
        if i < n - 1:
            h = torch.relu(h)
    return h


def evidence_activation(h: Tensor) -> Tensor:
    """``activation_function(h, 'exp')`` (utils.py:46-63), same op order."""
    h = h.clamp(-10, 10)
    log1e13 = 13 * torch.log(torch.tensor(10.0, dtype=h.dtype))
    numerator = h + log1e13
    denominator = torch.logaddexp(h, log1e13)
    return torch.exp(numerator - denominator)


def evidential_head(x: Tensor, weights, biases) -> Tensor:
    """``EvidentialNN.forward`` (models/classifiers.py:497-502) in eval mode."""
    return evidence_activation(mlp(x, weights, biases))


# --------------------------------------------------------------------------------------
# a14 / a15  fusion rules                                              utils.py:66-116
# --------------------------------------------------------------------------------------
def fuse(evid: Tensor, agg: str) -> Tensor:
    """evid [B,V,C] -> [B,C].  cml utils.py:66-68, avg :71-73, disentangled :76-78,
    joint :81-86 (shared_index 0, weight .5), dbf :88-116 (lambda 3)."""
    if agg == "cml":
        return evid.sum(dim=1)
    if agg == "avg":
        return evid.mean(dim=1)
    if agg == "disentangled":
        return evid[:, 1:, :].sum(dim=1)
    if agg == "joint":
        return 0.5 * evid[:, 0, :] + 0.5 * evid[:, 1:, :].sum(dim=1)
    if agg == "dbf":
        return discounted_belief_fusion(evid)
    raise ValueError(agg)


def discounted_belief_fusion(evid: Tensor, flambda: float = 3) -> Tensor:
    """utils.py:88-116 without the host-syncing assert (:111)."""
    C = evid.shape[-1]
    V = evid.shape[1]
    denom = (evid + 1).sum(dim=-1, keepdim=True)
    prob = (evid + 1) / denom
    belief = evid / denom
    unc = C / denom                                   # [B,V,1]
    discount = torch.ones(belief.shape[:-1], dtype=torch.float32).to(evid.dtype)
    for i in range(V):
        cp = torch.abs(prob[:, i].unsqueeze(1) - prob).sum(-1) / 2     # [B,V]
        cc = ((1 - unc[:, i].unsqueeze(1)) * (1 - unc)).squeeze(-1)     # [B,V]
        dc = cp * cc
        agreement = torch.prod((1 - dc ** flambda) ** (1 / flambda), dim=1)
        discount[:, i] = discount[:, i] * agreement
    discount = discount.unsqueeze(-1)
    belief = belief * discount
    unc = unc * discount + 1 - discount
    return (C * belief / (unc + 1e-6)).mean(dim=1)


# --------------------------------------------------------------------------------------
# a16  AvgTrustedLoss                                         models/losses.py:117-248
# --------------------------------------------------------------------------------------
def _dirichlet_kl_to_uniform(alpha: Tensor, C: int) -> Tensor:
    """models/losses.py:189-204."""
    ones = torch.ones([1, C], dtype=alpha.dtype)
    s = alpha.sum(dim=1, keepdim=True)
    first = (torch.lgamma(s) - torch.lgamma(alpha).sum(dim=1, keepdim=True)
             + torch.lgamma(ones).sum(dim=1, keepdim=True)
             - torch.lgamma(ones.sum(dim=1, keepdim=True)))
    second = ((alpha - ones) * (torch.digamma(alpha) - torch.digamma(s))).sum(dim=1, keepdim=True)
    return first + second


def edl_digamma_loss(alpha: Tensor, y1h: Tensor, annealing_step, C: int, annealing_start) -> Tensor:
    """models/losses.py:117-138: mean over rows of A + coef*KL."""
    S = alpha.sum(dim=1, keepdim=True)
    A = (y1h * (torch.digamma(S) - torch.digamma(alpha))).sum(dim=1, keepdim=True)
    coef = torch.min(torch.tensor(1.0, dtype=torch.float32),
                     torch.tensor(annealing_step / annealing_start, dtype=torch.float32)).to(alpha.dtype)
    kl_alpha = (alpha - 1) * (1 - y1h) + 1
    return (A + coef * _dirichlet_kl_to_uniform(kl_alpha, C)).mean()


def dc_loss(evid: Tensor, eps: float = 1e-8) -> Tensor:
    """``get_dc_loss_vectorized`` models/losses.py:161-187."""
    B, V, C = evid.shape
    alpha = evid + 1.0
    S = alpha.sum(dim=-1, keepdim=True)
    p = alpha / (S + eps)
    u = (C / (S + eps)).squeeze(-1)
    pd = (p.unsqueeze(2) - p.unsqueeze(1)).abs().sum(dim=-1) * 0.5
    om = 1.0 - u
    cc = om.unsqueeze(2) * om.unsqueeze(1)
    return ((pd * cc).sum(dim=2) / max(1, V - 1)).sum(dim=1).mean()


def avg_trusted_loss(evid: Tensor, target: Tensor, evidence_a: Tensor, fused: float,
                     annealing_step: int, annealing_start: int, gamma: float = 1.0) -> Tensor:
    """``AvgTrustedLoss.forward`` models/losses.py:217-248.  The fused-evidence EDL term is
    computed and dropped by the reference (:226-228 vs :239-240); so it is here."""
    B, V, C = evid.shape
    y1h = F.one_hot(target, C).to(evid.dtype)
    alpha_flat = (evid + 1).reshape(B * V, C)
    y_flat = y1h.repeat_interleave(V, dim=0)
    loss_views_mean = edl_digamma_loss(alpha_flat, y_flat, annealing_step, C, annealing_start)
    loss_acc = loss_views_mean / V
    t = min(1.0, annealing_step / max(1, annealing_start))
    gamma_t = 0.2 * (1 - t) + gamma * t
    return loss_acc + gamma_t * dc_loss(evid) * fused


# --------------------------------------------------------------------------------------
# a17  uncertainty summaries        models/evidential_probe.py:139-143, analysis.py:27-34
# --------------------------------------------------------------------------------------
def uncertainty_summaries(fused_evid: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """returns (epistemic u = C/S  [B], aleatoric [B], argmax [B])."""
    C = fused_evid.shape[-1]
    alphas = fused_evid + 1
    denom = alphas.sum(dim=-1, keepdim=True)
    probs = alphas / denom
    u = (C / denom).squeeze(-1)
    ale = -torch.sum(probs * (torch.digamma(alphas + 1) - torch.digamma(denom + 1)), dim=-1)
    return u, ale, fused_evid.argmax(dim=-1)


# --------------------------------------------------------------------------------------
# a9  SupConLoss (default args)                                 models/losses.py:17-101
# --------------------------------------------------------------------------------------
def supcon(z0: Tensor, z1: Tensor, T: float = 0.07) -> Tuple[Tensor, Tensor, Tensor]:
    """Restates the default path (contrast_mode='all', no labels/mask): features [B,2,D].
    Follows the reference's materialised form; only for sizes where [2B,2B] fits."""
    B = z0.shape[0]
    cf = torch.cat([z0, z1], dim=0)                                # :53
    adc = torch.matmul(cf, cf.T) / T                               # :64-66
    logits = adc - adc.max(dim=1, keepdim=True).values.detach()    # :68-69
    eye = torch.eye(B, dtype=z0.dtype, device=z0.device)
    mask = eye.repeat(2, 2)                                        # :72
    lmask = torch.ones_like(mask)
    lmask[:B, :B] = 0
    lmask[B:, B:] = 0
    mask = mask * lmask
    exp_logits = torch.exp(logits) * lmask                         # :79
    log_prob = logits - torch.log(exp_logits.sum(1, keepdim=True) + 1e-12)
    mlpp = (mask * log_prob).sum(1) / mask.sum(1)
    loss = (-mlpp).view(2, B).mean()                               # :86-87 (T/base_T = 1)
    with torch.no_grad():                                          # :89-99
        lmx = torch.ones_like(mask)
        lmx[:B, B:] = 0
        lmx[B:, :B] = 0
        elx = torch.exp(logits) * lmx
        lpx = logits - torch.log(elx.sum(1, keepdim=True))
        mx = torch.zeros_like(mask)
        mx.diagonal().fill_(1)
        m = (mx * lpx).sum(1) / mx.sum(1)
        lx, ly = (-m).view(2, B).mean(1)
    return loss, lx, ly


def supcon_stored_probabilities(z0: Tensor, z1: Tensor, T: float = 0.07):
    """The quantities of the stored-probability form of SupConLoss's backward for L2-NORMALISED views (what
    models/disentangledssl.py:134-140 feeds; models/losses.py:64-99 under autograd):
        e_ij = exp(s_ij - 1/T),  lseA_i = log sum_j exp(s_ij),  lseB_j = log sum_i exp(s_ij)   (cross block only: the
        reference masks the intra-view logits out of the denominator),
        exp(s_ij - lseA_i) + exp(s_ij - lseB_j) = e_ij (fa_i + fb_j),  fa = exp(1/T - lseA),  fb = exp(1/T - lseB),
        dz0 = (W z1 - 2 z1) / (2 B T),  dz1 = (W^T z0 - 2 z0) / (2 B T)  with  W = e (fa + fb).
    Returns (e, fa, fb, dz0, dz1)."""
    B = z0.shape[0]
    s = z0 @ z1.T / T
    e = torch.exp(s - 1.0 / T)
    fa, fb = 1.0 / e.sum(1), 1.0 / e.sum(0)
    W = e * (fa[:, None] + fb[None, :])
    coef = 1.0 / (2 * B * T)
    return e, fa, fb, coef * (W @ z1 - 2 * z1), coef * (W.T @ z0 - 2 * z0)


def infonce_e_blocks(e: Tensor) -> Tensor:
    """Dense [Ma, Nb] probabilities -> the blocked HBM layout of dmf_infonce_rowcol_sums_store / dmf_infonce_bwd_stored
    (include/dmf_b200.h): blocks of [128 rows x 64 columns], block (ib, jb) at index ib * njb + jb, rows padded to a
    multiple of 256 and columns to a multiple of 256 with zeros.  Returns [nib * njb, 128, 64]."""
    Ma, Nb = e.shape
    nib, njb = 2 * ((Ma + 255) // 256), 4 * ((Nb + 255) // 256)
    pad = torch.zeros(nib * 128, njb * 64, dtype=e.dtype)
    pad[:Ma, :Nb] = e
    return pad.view(nib, 128, njb, 64).permute(0, 2, 1, 3).reshape(nib * njb, 128, 64).contiguous()


def infonce_e_unblock(blocks: Tensor, Ma: int, Nb: int) -> Tensor:
    """inverse of ``infonce_e_blocks``"""
    nib, njb = 2 * ((Ma + 255) // 256), 4 * ((Nb + 255) // 256)
    return blocks.view(nib, njb, 128, 64).permute(0, 2, 1, 3).reshape(nib * 128, njb * 64)[:Ma, :Nb]


def ortho_loss(z1: Tensor, zs: Tensor) -> Tensor:
    """models/losses.py:104-110."""
    return torch.norm(torch.matmul(F.normalize(z1, dim=-1).T, F.normalize(zs, dim=-1)))


# --------------------------------------------------------------------------------------
# a8  vMF reparameterised sample given explicit noise     models/classifiers.py:314-335,433-466
# --------------------------------------------------------------------------------------
def vmf_rsample(e: Tensor, w: Tensor, v: Tensor) -> Tensor:
    """``ProbabilisticEncoder('vmf')`` + ``VonMisesFisher.rsample`` with the random parts
    supplied: ``w`` [B,1] is the rejection-sampled first coordinate, ``v`` [B,D-1] the
    *normalised* tangent direction.  loc = e/||e|| (:463); x=[w, sqrt(clamp(1-w^2,1e-10)) v]
    (:331-332); Householder about u=(e1-loc)/(||e1-loc||+1e-5) (:433-437)."""
    loc = e / e.norm(dim=-1, keepdim=True)
    w_ = torch.sqrt(torch.clamp(1 - w ** 2, 1e-10))
    x = torch.cat((w, w_ * v), -1)
    e1 = torch.zeros(e.shape[-1], dtype=e.dtype, device=e.device)
    e1[0] = 1.0
    u = e1 - loc
    u = u / (u.norm(dim=-1, keepdim=True) + 1e-5)
    return x - 2 * (x * u).sum(-1, keepdim=True) * u


def draw_vmf_noise(B: int, D: int, kappa: float = 1.0, k: int = 1,
                   dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """Replays the reference's CPU draw order for one ``rsample()`` call
    (models/classifiers.py:314-431): rejection loop {Beta(fp64) [B,k], Uniform [B,k]} until
    every row accepted, then Normal [B,D] whose first column is dropped and rows normalised.
    Consumes the global CPU torch generator exactly like the reference, so after
    ``torch.manual_seed(s)`` the returned (w, v) are the reference's own."""
    m = D
    scale = kappa * torch.ones(B, 1, dtype=dtype)
    c = torch.sqrt((4 * (scale ** 2)) + (m - 1) ** 2)
    b_true = (-2 * scale + c) / (m - 1)
    b_app = (m - 1) / (4 * scale)
    s = torch.min(torch.max(torch.tensor([0.0], dtype=dtype), scale - 10),
                  torch.tensor([1.0], dtype=dtype))
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
        w_sel = w_.gather(1, idx)
        take = active & any_acc
        w[take] = w_sel[take]
        active = active & ~any_acc
    v = torch.distributions.Normal(0, 1).sample(torch.Size([B, D])).type(dtype)[:, 1:]
    v = v / v.norm(dim=-1, keepdim=True)
    return w, v


# --------------------------------------------------------------------------------------
# a3-a6, a11  DMVAE                                              models/dmvae.py:74-188
# --------------------------------------------------------------------------------------
def _gauss_kl(mu: Tensor, logvar: Tensor) -> Tensor:
    """models/dmvae.py:86-89."""
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1)


def product_of_experts(mu_list, logvar_list, temperature: float) -> Tuple[Tensor, Tensor]:
    """models/dmvae.py:91-112 with include_prior=True."""
    mus = torch.stack(list(mu_list) + [torch.zeros_like(mu_list[0])], dim=0)
    lvs = torch.stack(list(logvar_list) + [torch.zeros_like(logvar_list[0])], dim=0)
    prec = torch.exp(-lvs) / max(temperature, 1e-8)
    psum = prec.sum(dim=0) + 1e-8
    var = 1.0 / psum
    mu = var * (prec * mus).sum(dim=0)
    return mu, torch.log(var)


def dmvae_forward(xs: Sequence[Tensor], enc: Sequence[Tuple[list, list]],
                  dec: Sequence[Tuple[list, list]], noise: Sequence[Tensor], a: float,
                  cross_weight: float = 1.0, lam: Optional[Sequence[float]] = None
                  ) -> Tuple[Tensor, Dict[str, Tensor]]:
    """``DMVAE.forward`` models/dmvae.py:128-188.  ``noise`` is the 2N+1 ``randn_like`` draws in
    the reference's order (:147-150): z_p views 0..N-1, z_s_uni views 0..N-1, then z_s (PoE).
    PoE temperature is the hard-coded 1.5 of :149."""
    N = len(xs)
    lam = list(lam) if lam is not None else [1.0] * N
    stats = [mlp(xs[i], *enc[i]).chunk(4, dim=1) for i in range(N)]    # [mu_s, lv_s, mu_p, lv_p]
    mu_s = [s[0] for s in stats]
    lv_s = [s[1] for s in stats]
    mu_p = [s[2] for s in stats]
    lv_p = [s[3] for s in stats]
    z_p = [mu_p[i] + noise[i] * torch.exp(0.5 * lv_p[i]) for i in range(N)]
    z_su = [mu_s[i] + noise[N + i] * torch.exp(0.5 * lv_s[i]) for i in range(N)]
    mu_poe, lv_poe = product_of_experts(mu_s, lv_s, 1.5)
    z_s = mu_poe + noise[2 * N] * torch.exp(0.5 * lv_poe)
    feats = [x.to(z_s.dtype) for x in xs]
    rec_joint = sum(lam[i] * F.mse_loss(mlp(torch.cat([z_p[i], z_s], 1), *dec[i]), feats[i])
                    for i in range(N))
    rec_cross = 0.0
    pairs = 0
    for i in range(N):
        for j in range(N):
            if i == j:
                continue
            rec_cross = rec_cross + lam[i] * F.mse_loss(
                mlp(torch.cat([z_p[i], z_su[j]], 1), *dec[i]), feats[i])
            pairs += 1
    rec_cross = rec_cross / pairs * cross_weight
    kl_private = torch.stack([_gauss_kl(mu_p[i], lv_p[i]) for i in range(N)], 1).sum(1).mean()
    kl_poe = _gauss_kl(mu_poe, lv_poe).mean()
    kl_uni = torch.stack([_gauss_kl(mu_s[i], lv_s[i]) for i in range(N)], 1).sum(1).mean()
    loss = rec_joint + a * (kl_private + N * kl_poe) + rec_cross + a * kl_uni
    logs = {"loss": loss.detach(), "loss_joint_recon": rec_joint.detach(),
            "loss_cross_recon": rec_cross.detach(), "kl_private": kl_private.detach(),
            "kl_shared_poe": kl_poe.detach(), "kl_shared_uni_sum": kl_uni.detach()}
    return loss, logs


def dmvae_get_embedding(xs, enc, poe_temperature: float = 1.5):
    """``DMVAE.get_embedding(return_poe=True)`` models/dmvae.py:115-125."""
    stats = [mlp(xs[i], *enc[i]).chunk(4, dim=1) for i in range(len(xs))]
    mu, _ = product_of_experts([s[0] for s in stats], [s[1] for s in stats], poe_temperature)
    return mu, [s[2] for s in stats]


# --------------------------------------------------------------------------------------
# a7  DisentangledSSL                                  models/disentangledssl.py:67-160
# --------------------------------------------------------------------------------------
def dssl_forward(x1, x2, v1, v2, p: Dict[str, Tuple[list, list]], noise: Sequence,
                 a: float = 1.0, lmd: float = 0.0, T: float = 0.07, condzs: bool = True, usezsx: bool = False,
                 distribution: str = "vmf") -> Tuple[Tensor, Dict[str, Tensor]]:
    """``DisentangledSSL.forward`` models/disentangledssl.py:82-160.
    ``p`` maps 'x1s','x2s','x1','x2' -> (weights, biases) of encoder_x1s/x2s/x1/x2.
    ``noise`` in the reference's rsample order (:100-103: zs1, zs2, zsv1, zsv2): four (w, v) pairs for
    distribution='vmf', four standard-normal tensors [B, D] for 'normal' (Independent(Normal(mu, 1)).rsample() =
    mu + eps, models/classifiers.py:456-459).  ``condzs`` (:116-125): private encoders see [x | e] or x alone;
    ``usezsx`` (:128-137): the specific critic compares normalize([z | e]) instead of normalize(z)."""
    e1, e2 = mlp(x1, *p["x1s"]), mlp(x2, *p["x2s"])                # :90-93
    e1v, e2v = mlp(v1, *p["x1s"]), mlp(v2, *p["x2s"])
    if distribution == "vmf":
        zs1, zs2 = vmf_rsample(e1, *noise[0]), vmf_rsample(e2, *noise[1])
        zsv1, zsv2 = vmf_rsample(e1v, *noise[2]), vmf_rsample(e2v, *noise[3])
    elif distribution == "normal":
        zs1, zs2, zsv1, zsv2 = e1 + noise[0], e2 + noise[1], e1v + noise[2], e2v + noise[3]
    else:
        raise ValueError(distribution)
    j, lx, ly = supcon(zs1, zs2, T)                                # :106-113
    jv, lxv, lyv = supcon(zsv1, zsv2, T)
    joint = 0.5 * (j + jv)
    loss_x, loss_y = 0.5 * (lx + lxv), 0.5 * (ly + lyv)
    if condzs:                                                     # :116-125
        z1x1 = mlp(torch.cat([x1, e1], 1), *p["x1"])
        z1xv1 = mlp(torch.cat([v1, e1v], 1), *p["x1"])
        z2x2 = mlp(torch.cat([x2, e2], 1), *p["x2"])
        z2xv2 = mlp(torch.cat([v2, e2v], 1), *p["x2"])
    else:
        z1x1, z1xv1 = mlp(x1, *p["x1"]), mlp(v1, *p["x1"])
        z2x2, z2xv2 = mlp(x2, *p["x2"]), mlp(v2, *p["x2"])
    n = lambda t: F.normalize(t, dim=-1)                           # :128-143
    if usezsx:
        c1 = (n(torch.cat([z1x1, e1], 1)), n(torch.cat([z1xv1, e1v], 1)))
        c2 = (n(torch.cat([z2x2, e2], 1)), n(torch.cat([z2xv2, e2v], 1)))
    else:
        c1, c2 = (n(z1x1), n(z1xv1)), (n(z2x2), n(z2xv2))
    s1, _, _ = supcon(*c1, T)
    s2, _, _ = supcon(*c2, T)
    specific = s1 + s2
    ortho = 0.5 * (ortho_loss(z1x1, e1) + ortho_loss(z2x2, e2)) + \
        0.5 * (ortho_loss(z1xv1, e1v) + ortho_loss(z2xv2, e2v))    # :154-155
    loss = 2 * joint / (1 + a) + a * specific / (1 + a) + lmd * ortho   # :157
    logs = {"loss": loss.detach(), "shared": joint.detach(), "clip": joint.detach(),
            "loss_x": loss_x, "loss_y": loss_y, "specific": specific.detach(),
            "ortho": ortho.detach()}
    return loss, logs


def dssl_get_embedding(x1, x2, p, condzs: bool = True):
    """``DisentangledSSL.get_embedding`` models/disentangledssl.py:67-80."""
    zs1, zs2 = mlp(x1, *p["x1s"]), mlp(x2, *p["x2s"])
    if condzs:
        z1 = mlp(torch.cat([x1, zs1], 1), *p["x1"])
        z2 = mlp(torch.cat([x2, zs2], 1), *p["x2"])
    else:
        z1, z2 = mlp(x1, *p["x1"]), mlp(x2, *p["x2"])
    return torch.cat([zs1, zs2], 1), [z1, z2]


# --------------------------------------------------------------------------------------
# a13  probe shared_step                 models/evidential_probe.py:87-103, baselines.py:42-70
# --------------------------------------------------------------------------------------
def probe_shared_step(embeds: Sequence[Tensor], heads: Sequence[Tuple[list, list]], target: Tensor,
                      agg: str, fused: float, annealing_step: int, annealing_start: int):
    """heads applied to ``embeds`` one-to-one -> stack [B,V,C] -> agg -> AvgTrustedLoss.
    Returns the reference's 4-tuple (loss, evidences_a, target, evidences)."""
    evid = torch.stack([evidential_head(embeds[i], *heads[i]) for i in range(len(heads))], dim=1)
    ea = fuse(evid, agg)
    loss = avg_trusted_loss(evid, target, ea, fused, annealing_step, annealing_start)
    return loss, ea, target, evid


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench baselines
# --------------------------------------------------------------------------------------
def xavier_mlp_params(layers: Sequence[int], out: int, gen: torch.Generator, dtype=torch.float32):
    """Random-init (weights, biases) with the reference's init family (xavier_uniform weights,
    nn.Linear-default biases; utils.py:153-157).  Not seed-equal to the reference; used for
    synthetic benchmarks only."""
    dims = list(layers) + [out]
    ws, bs = [], []
    for i in range(len(dims) - 1):
        fan_in, fan_out = dims[i], dims[i + 1]
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        ws.append(((torch.rand(fan_out, fan_in, generator=gen, dtype=dtype) * 2 - 1) * bound).requires_grad_())
        bb = 1.0 / math.sqrt(fan_in)
        bs.append(((torch.rand(fan_out, generator=gen, dtype=dtype) * 2 - 1) * bb).requires_grad_())
    return ws, bs


# --------------------------------------------------------------------------------------
# evaluation reducer                                                     analysis.py:5-192
# --------------------------------------------------------------------------------------
def eval_reduce(evidences: Tensor, fused_ev: Tensor, target: Tensor) -> Dict[str, Tensor]:
    """Accumulators of ``evaluate_subjective_model`` for ONE batch (analysis.py:27-152), slot V = fused:
    stats [V+1, 8] = correct, evidence_sum, epi_sum, ale_sum, inc_N, inc_evidence_sum, inc_epi_sum, inc_ale_sum;
    class_sum / true_sum [V+1, K]; class_counts [K]."""
    B, V, K = evidences.shape
    slots = [evidences[:, v, :] for v in range(V)] + [fused_ev]
    stats = torch.zeros(V + 1, 8, dtype=torch.float64)
    class_sum = torch.zeros(V + 1, K, dtype=torch.float64)
    true_sum = torch.zeros(V + 1, K, dtype=torch.float64)
    for s, ev in enumerate(slots):
        alphas = ev + 1.0
        S = alphas.sum(dim=-1, keepdim=True)
        probs = alphas / S
        epi = (K / S).squeeze(-1)
        ale = -torch.sum(probs * (torch.digamma(alphas + 1.0) - torch.digamma(S + 1.0)), dim=-1)
        escal = ev.sum(dim=-1)
        ok = ev.argmax(dim=-1) == target
        inc = ~ok
        stats[s] = torch.tensor([ok.sum(), escal.sum(), epi.sum(), ale.sum(), inc.sum(), escal[inc].sum(), epi[inc].sum(),
                                 ale[inc].sum()], dtype=torch.float64)
        class_sum[s] = ev.sum(dim=0).double()
        true_sum[s] = torch.bincount(target, weights=ev[torch.arange(B), target], minlength=K).double()
    return dict(stats=stats, class_sum=class_sum, true_sum=true_sum, class_counts=torch.bincount(target, minlength=K).double())
