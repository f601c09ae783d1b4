"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests / golden vectors of its own, so these fixtures --
inputs, weights, explicit noise, outputs and gradients of the reference modules on seeded
inputs, true-fp32 ('highest' matmul precision, CPU) -- are what pins ``oracle/port.py`` and,
through it and directly, the CUDA path.  Everything is stored as float32/int64 ``.npz``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import port  # noqa: E402
from oracle.ref_harness import cuda_identity_shim, in_reference_cwd, load_reference  # noqa: E402

ns = load_reference()
torch.set_num_threads(1)          # deterministic reduction order in the fixtures


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (npy(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


def state_arrays(module, prefix="sd."):
    return {prefix + k: v for k, v in module.state_dict().items()}


# ----------------------------------------------------------------------------------------
# 1. evidence activation                                                   utils.py:46-63
# ----------------------------------------------------------------------------------------
def gen_activation():
    g = torch.Generator().manual_seed(11)
    h = torch.cat([torch.linspace(-12, 12, 97), torch.randn(927, generator=g) * 4,
                   torch.tensor([-10.0, 10.0, 0.0, -0.0, 9.999999, -9.999999])])
    h = h.requires_grad_()
    e = ns.utils.activation_function(h, "exp")
    (gr,) = torch.autograd.grad(e.sum(), h)
    save("activation", h=h, e=e, grad=gr)


# ----------------------------------------------------------------------------------------
# 2. fusion rules + AvgTrustedLoss + uncertainty summaries
#    utils.py:66-116, models/losses.py:117-248, models/evidential_probe.py:139-143
# ----------------------------------------------------------------------------------------
def gen_edl():
    cfgs = {"c1": (100, 7, 10, 50), "c2": (128, 3, 3, 10), "c3scene": (100, 4, 15, 50),
            "c3late": (100, 3, 15, 50), "c4": (64, 4, 42, 20), "ragged": (37, 2, 5, 20),
            "single": (1, 3, 4, 20)}
    for tag, (B, V, C, astart) in cfgs.items():
        g = torch.Generator().manual_seed(sum(map(ord, tag)) + 5)
        h = (torch.randn(B, V, C, generator=g) * 2.0).clamp(-10, 10)
        evid = ns.utils.activation_function(h, "exp")
        y = torch.randint(0, C, (B,), generator=g)
        out = {"evid": evid, "y": y, "annealing_start": astart}
        for agg, fn in (("cml", ns.utils.get_cml_fusion), ("avg", ns.utils.get_avg_fusion),
                        ("joint", ns.utils.get_joint_fusion),
                        ("disentangled", ns.utils.get_disentangled_fusion),
                        ("dbf", ns.utils.discounted_belief_fusion)):
            fe = fn(evid.clone())
            out["fused_" + agg] = fe
            alphas = fe + 1                                     # evidential_probe.py:139-143
            denom = alphas.sum(dim=-1, keepdim=True)
            probs = alphas / denom
            out["u_" + agg] = (C / denom).squeeze(-1)
            out["ale_" + agg] = -torch.sum(probs * (torch.digamma(alphas + 1) - torch.digamma(denom + 1)), dim=-1)
            out["pred_" + agg] = fe.argmax(dim=-1)
        out["pred_views"] = evid.argmax(dim=-1)
        for step in (0, 7, astart + 10):
            for fused in (1, 0):
                crit = ns.losses.AvgTrustedLoss(num_views=V, annealing_start=astart)
                crit.annealing_step = step
                ev = evid.clone().requires_grad_()
                loss = crit(ev, y, ns.utils.get_cml_fusion(ev), fused=fused)
                (gr,) = torch.autograd.grad(loss, ev)
                out[f"loss_s{step}_f{fused}"] = loss
                out[f"grad_s{step}_f{fused}"] = gr
        out["steps"] = np.array([0, 7, astart + 10])
        save("edl_" + tag, **out)


# ----------------------------------------------------------------------------------------
# 3. SupConLoss / ortho_loss                                    models/losses.py:17-110
# ----------------------------------------------------------------------------------------
def gen_supcon():
    for tag, (B, D, unit) in {"b64d16": (64, 16, True), "b96d64": (96, 64, True),
                              "b33d24raw": (33, 24, False), "b2d8": (2, 8, True)}.items():
        g = torch.Generator().manual_seed(B * 7 + D)
        z0 = torch.randn(B, D, generator=g)
        z1 = (0.6 * z0 + 0.8 * torch.randn(B, D, generator=g))
        if unit:
            z0 = z0 / z0.norm(dim=-1, keepdim=True)
            z1 = z1 / z1.norm(dim=-1, keepdim=True)
        else:
            z0, z1 = 0.3 * z0, 0.3 * z1
        z0.requires_grad_()
        z1.requires_grad_()
        crit = ns.losses.SupConLoss()
        loss, lx, ly = crit(torch.stack([z0, z1], dim=1))
        g0, g1 = torch.autograd.grad(loss, (z0, z1))
        a = torch.randn(B, D, generator=g).requires_grad_()
        b = torch.randn(B, D, generator=g).requires_grad_()
        ol = ns.losses.ortho_loss(a, b)
        ga, gb = torch.autograd.grad(ol, (a, b))
        save("supcon_" + tag, z0=z0, z1=z1, loss=loss, loss_x=lx, loss_y=ly, g0=g0, g1=g1,
             oa=a, ob=b, ortho=ol, goa=ga, gob=gb)


# ----------------------------------------------------------------------------------------
# 4. DMVAE                                                            models/dmvae.py
# ----------------------------------------------------------------------------------------
class _RecordRandnLike:
    def __enter__(self):
        self.rec = []
        self._orig = torch.randn_like

        def rl(t, *a, **k):
            n = self._orig(t, *a, **k)
            self.rec.append(n.clone())
            return n
        torch.randn_like = rl
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig
        return False


def gen_dmvae():
    for tag, (dims, h, e, B, a) in {"scene_small": ([20, 59, 40], 32, 8, 16, 1e-5),
                                    "hw_small": ([240, 76, 216, 47, 64, 6], 48, 12, 10, 1e-5),
                                    "syn_small": ([32, 32], 64, 16, 40, 1.0)}.items():
        torch.manual_seed(3)
        m = ns.dmvae.DMVAE(output_dim=dims, a=a, hidden_dim=h, embed_dim=e)
        g = torch.Generator().manual_seed(17)
        xs = [torch.rand(B, d, generator=g) for d in dims]
        torch.manual_seed(99)
        with _RecordRandnLike() as rr:
            loss, logs = m(xs)
        loss.backward()
        out = state_arrays(m)
        out.update({f"grad.{k}": p.grad for k, p in m.named_parameters()})
        out.update({f"x{i}": x for i, x in enumerate(xs)})
        out.update({f"noise{i}": n for i, n in enumerate(rr.rec)})
        out["loss"] = loss
        for k in ("loss_joint_recon", "loss_cross_recon", "kl_private", "kl_shared_poe", "kl_shared_uni_sum"):
            out["log." + k] = np.float32(logs[k])
        mu_poe, mu_p = m.get_embedding(xs)
        out["emb_shared"] = mu_poe
        out.update({f"emb_private{i}": t for i, t in enumerate(mu_p)})
        out["meta"] = np.array([h, e, B], dtype=np.int64)
        out["dims"] = np.array(dims, dtype=np.int64)
        out["a"] = np.float64(a)
        save("dmvae_" + tag, **out)


# ----------------------------------------------------------------------------------------
# 5. DisentangledSSL                                        models/disentangledssl.py
# ----------------------------------------------------------------------------------------
def gen_dssl():
    for tag, (dims, h, e, B, a) in {"small": ([24, 40], 32, 16, 32, 1.0),
                                    "wide": ([72, 64], 64, 32, 48, 0.5)}.items():
        with cuda_identity_shim():
            torch.manual_seed(5)
            m = ns.disentangledssl.DisentangledSSL(output_dim=dims, hidden_dim=h, embed_dim=e, a=a,
                                                   lmd_start_value=0.25)
            g = torch.Generator().manual_seed(23)
            x1, x2 = torch.randn(B, dims[0], generator=g), torch.randn(B, dims[1], generator=g)
            v1 = x1 + 0.01 * torch.randn(B, dims[0], generator=g)
            v2 = x2 + 0.01 * torch.randn(B, dims[1], generator=g)
            torch.manual_seed(1234)
            loss, logs = m(x1, x2, v1, v2)
            loss.backward()
            # replay the reference's draw order to obtain the explicit noise
            torch.manual_seed(1234)
            noise = [port.draw_vmf_noise(B, e, 1.0) for _ in range(4)]
            emb_s, emb_p = m.get_embedding([x1, x2])
        out = state_arrays(m)
        out.update({f"grad.{k}": p.grad for k, p in m.named_parameters()})
        out.update(x1=x1, x2=x2, v1=v1, v2=v2, loss=loss, emb_shared=emb_s,
                   emb_private0=emb_p[0], emb_private1=emb_p[1])
        for i, (w, v) in enumerate(noise):
            out[f"noise_w{i}"] = w
            out[f"noise_v{i}"] = v
        for k in ("shared", "clip", "loss_x", "loss_y", "specific", "ortho", "lmd"):
            out["log." + k] = np.float32(logs[k])
        out["meta"] = np.array([h, e, B], dtype=np.int64)
        out["dims"] = np.array(dims, dtype=np.int64)
        out["a"] = np.float64(a)
        save("dssl_" + tag, **out)


# ----------------------------------------------------------------------------------------
# 6. probes / LateFusion (eval mode: dropout off)
#    models/evidential_probe.py:87-103,289-304, models/baselines.py:42-70
# ----------------------------------------------------------------------------------------
def gen_probes():
    dims, h, e, B, C = [20, 59, 40], 32, 8, 24, 15
    torch.manual_seed(3)
    backbone = ns.dmvae.DMVAE(output_dim=dims, a=1e-5, hidden_dim=h, embed_dim=e)
    g = torch.Generator().manual_seed(41)
    xs = [torch.rand(B, d, generator=g) for d in dims]
    y = torch.randint(0, C, (B,), generator=g)
    base = state_arrays(backbone, "backbone_sd.")
    base.update({f"x{i}": x for i, x in enumerate(xs)})
    base["y"] = y
    base["dims"] = np.array(dims, dtype=np.int64)
    base["meta"] = np.array([h, e, B, C], dtype=np.int64)

    for agg in ("cml", "avg", "joint", "disentangled"):
        torch.manual_seed(8)
        pm = ns.evidential_probe.EvidentialProbeModule(backbone, num_classes=C, input_dim=e,
                                                       hidden_dim=(16,), dropout=0.1, annealing_start=50,
                                                       aggregation=agg, fused=1)
        pm.eval()
        pm.criterion.annealing_step = 5
        loss, ea, _, ev = pm.shared_step([*xs, y])
        loss.backward()
        out = dict(base)
        out.update(state_arrays(pm))
        out.update({f"grad.{k}": p.grad for k, p in pm.named_parameters() if p.grad is not None})
        out.update(loss=loss, evidences_a=ea, evidences=ev, annealing_step=5, annealing_start=50)
        save("probe_" + agg, **out)

    torch.manual_seed(8)
    dm = ns.evidential_probe.DisentangledEvidentialProbeModule(backbone, num_classes=C, input_dim=e,
                                                               hidden_dim=(16,), dropout=0.1,
                                                               annealing_start=50, aggregation="cml")
    dm.eval()
    dm.criterion.annealing_step = 60
    loss, ea, _, ev = dm.shared_step([*xs, y])
    loss.backward()
    out = dict(base)
    out.update(state_arrays(dm))
    out.update({f"grad.{k}": p.grad for k, p in dm.named_parameters() if p.grad is not None})
    out.update(loss=loss, evidences_a=ea, evidences=ev, annealing_step=60, annealing_start=50)
    save("probe_dis_cml", **out)

    for agg in ("dbf", "cml", "avg"):
        torch.manual_seed(8)
        lf = ns.baselines.LateFusion([(ns.classifiers.IdentityEncoder, {}) for _ in dims], dims, C,
                                     dropout=0.1, aggregation=agg, annealing_start=50, hidden_dim=(16,))
        lf.eval()
        lf.criterion.annealing_step = 20
        loss, ea, _, ev = lf.shared_step([*xs, y])
        loss.backward()
        out = {f"x{i}": x for i, x in enumerate(xs)}
        out.update(y=y, dims=np.array(dims, dtype=np.int64), meta=np.array([0, 0, B, C], dtype=np.int64))
        out.update(state_arrays(lf))
        out.update({f"grad.{k}": p.grad for k, p in lf.named_parameters() if p.grad is not None})
        out.update(loss=loss, evidences_a=ea, evidences=ev, annealing_step=20, annealing_start=50)
        save("latefusion_" + agg, **out)


# ----------------------------------------------------------------------------------------
# 7. real data slice: HandWritten (6 views) first rows after the dataset's MinMax scaling
#    datasets/dataset.py:164-221,273-279  -> LateFusion(cml) on the real features
# ----------------------------------------------------------------------------------------
def gen_handwritten():
    with in_reference_cwd():
        ds = ns.dataset.HandWritten()
    idx = np.arange(0, 2000, 31)[:64]
    xs = [torch.from_numpy(np.stack([ds[i][v] for i in idx])) for v in range(ds.num_views)]
    y = torch.from_numpy(np.array([ds[i][-1] for i in idx], dtype=np.int64))
    dims = [int(d) for d in np.squeeze(ds.dims)]
    torch.manual_seed(8)
    lf = ns.baselines.LateFusion([(ns.classifiers.IdentityEncoder, {}) for _ in dims], dims, ds.num_classes,
                                 dropout=0.1, aggregation="cml", annealing_start=50, hidden_dim=(32,))
    lf.eval()
    lf.criterion.annealing_step = 3
    loss, ea, _, ev = lf.shared_step([*xs, y])
    loss.backward()
    out = {f"x{i}": x for i, x in enumerate(xs)}
    out.update(y=y, dims=np.array(dims, dtype=np.int64), meta=np.array([0, 0, len(idx), ds.num_classes], dtype=np.int64))
    out.update(state_arrays(lf))
    out.update({f"grad.{k}": p.grad for k, p in lf.named_parameters() if p.grad is not None})
    out.update(loss=loss, evidences_a=ea, evidences=ev, annealing_step=3, annealing_start=50)
    save("latefusion_handwritten", **out)


# ----------------------------------------------------------------------------------------
# 8. vMF sampler draw order                              models/classifiers.py:314-431
# ----------------------------------------------------------------------------------------
def gen_vmf():
    B, D = 40, 16
    g = torch.Generator().manual_seed(2)
    e = torch.randn(B, D, generator=g).requires_grad_()
    with cuda_identity_shim():
        ph = ns.classifiers.ProbabilisticEncoder(torch.nn.Identity(), distribution="vmf", vmfkappa=1)
        torch.manual_seed(77)
        dist, _ = ph(e)
        z = dist.rsample()
    (ge,) = torch.autograd.grad((z * torch.arange(D, dtype=torch.float32)).sum(), e)
    torch.manual_seed(77)
    w, v = port.draw_vmf_noise(B, D, 1.0)
    save("vmf", e=e, z=z, w=w, v=v, grad_e=ge)


# ----------------------------------------------------------------------------------------
# 9. evaluation reducers                                                 analysis.py:5-399
# ----------------------------------------------------------------------------------------
def gen_eval():
    """The reference's evaluate_subjective_model[_with_shared] on a stub model whose shared_step replays fixed
    evidences over three ragged batches; the returned dicts are stored flattened (JSON)."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("ref_analysis", os.path.join(os.environ.get("DMF_REFERENCE_ROOT", "/root/reference"), "analysis.py"))
    ana = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ana)
    g = torch.Generator().manual_seed(41)
    V, K = 4, 7
    sizes = [50, 33, 17]
    evs = [port.evidence_activation(torch.randn(b, V, K, generator=g) * 1.5) for b in sizes]
    ys = [torch.randint(0, K, (b,), generator=g) for b in sizes]
    fused = [e.sum(dim=1) for e in evs]

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.num_classes = K
            self.i = 0

        def shared_step(self, batch):
            i = self.i
            self.i += 1
            return torch.zeros(()), fused[i], ys[i], evs[i]
    loader = [[torch.zeros(b, 1), ys[i]] for i, b in enumerate(sizes)]
    r1 = ana.evaluate_subjective_model(Stub(), loader, device=torch.device("cpu"))
    r2 = ana.evaluate_subjective_model_with_shared(Stub(), loader, device=torch.device("cpu"))
    save("eval_reduce", evid=torch.cat(evs), y=torch.cat(ys), fused=torch.cat(fused), sizes=np.asarray(sizes),
         plain_json=np.frombuffer(json.dumps(r1).encode(), dtype=np.uint8),
         shared_json=np.frombuffer(json.dumps(r2).encode(), dtype=np.uint8))


# ----------------------------------------------------------------------------------------
# 10. datasets: synthetic generator + MultiViewDataset post-processing   datasets/dataset.py:164-268,331-458
# ----------------------------------------------------------------------------------------
def gen_datasets():
    import hashlib
    ds_mod = ns.dataset
    rows = {}
    for tag, kw in (("common_med", dict(n_samples=10000, n_classes=3, d_signal=16, d_spurious=16, rho=0.5, shared_class_frac=0.5, seed=0)),
                    ("small_linear", dict(n_samples=257, n_classes=4, d_signal=8, d_spurious=0, rho=0.9, shared_class_frac=0.25,
                                          hetero_noise=False, nonlinear_shared=False, nonlinear_specific=True, conflict_frac=1.0, seed=3))):
        d = ds_mod.SimpleTwoModalPlus(**kw)
        for nm, t in (("X1", d.X1), ("X2", d.X2), ("y", d.y)):
            rows[f"{tag}.{nm}.sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest().encode(), dtype=np.uint8)
        rows[f"{tag}.X1.head"] = d.X1[:4]
    # MultiViewDataset on seeded raw arrays (no .mat needed): normalisation, conflict + noise injection
    rng = np.random.RandomState(5)
    N, dims, Cn = 120, [7, 3, 5], 4
    raw = np.empty(len(dims), dtype=object)
    for v, dv in enumerate(dims):
        raw[v] = rng.randn(N, dv) * (v + 1) + v
    yy = rng.randint(1, Cn + 1, size=(N, 1))                     # 1-based labels like the .mat files
    mv = ds_mod.MultiViewDataset("toy", raw.copy(), yy.copy())
    np.random.seed(123)
    idx = np.arange(N)
    np.random.shuffle(idx)
    test_idx = idx[96:]
    mv.postprocessing(test_idx, addNoise=True, sigma=0.5, ratio_noise=0.5, addConflict=True, ratio_conflict=1.0)
    rows.update({"mv.raw0": raw[0], "mv.raw1": raw[1], "mv.raw2": raw[2], "mv.y_raw": yy, "mv.test_idx": test_idx,
                 "mv.X0": mv.X[0], "mv.X1": mv.X[1], "mv.X2": mv.X[2], "mv.Y": mv.Y, "mv.dims": mv.dims})
    save("datasets", **rows)


# ----------------------------------------------------------------------------------------
# 11. DMVAE at the BASELINE.json sizes (C1 HandWritten h=512 e=200 B=100 on real rows, C3 CUB 1024/300, C2 synthetic
#     B=4096).  The weights are NOT stored (C1 alone has 9.4 M parameters): the CUDA mirror's initialisation is
#     seed-equal to the reference's (tests/test_cpu_host.py::test_dmvae_init_stream_and_keys), so the fixture keeps the
#     seed, the inputs, the explicit noise, the outputs and -- per parameter -- the gradient norm plus 64 sampled
#     entries at fixed indices.                                          models/dmvae.py:128-188, run.py:179-208
# ----------------------------------------------------------------------------------------
def grad_samples(n, k=64):
    g = np.random.RandomState(n % (2 ** 31 - 1))
    return np.sort(g.choice(n, size=min(k, n), replace=False)).astype(np.int64)


def gen_dmvae_full():
    with in_reference_cwd():
        hw = ns.dataset.HandWritten()
        cub = ns.dataset.CUB()
    syn = ns.dataset.SimpleTwoModalPlus(n_samples=10000, n_classes=3, d_signal=16, d_spurious=16, rho=0.5,
                                         shared_class_frac=0.5, seed=0)

    def rows_of(ds, idx):
        return [torch.from_numpy(np.stack([ds[i][v] for i in idx])).float() for v in range(ds.num_views)]
    cases = {
        "c1_hw": (rows_of(hw, np.arange(0, 2000, 20)[:100]), 512, 200, 1e-5),
        "c3_cub": (rows_of(cub, np.arange(0, 600, 6)[:100]), 512, 200, 1e-5),
        "c2_syn": ([syn.X1[:4096].float(), syn.X2[:4096].float()], 512, 16, 1e-5),
    }
    for tag, (xs, h, e, a) in cases.items():
        dims = [int(x.shape[1]) for x in xs]
        B = xs[0].shape[0]
        torch.manual_seed(3)
        m = ns.dmvae.DMVAE(output_dim=dims, a=a, hidden_dim=h, embed_dim=e)
        torch.manual_seed(99)
        with _RecordRandnLike() as rr:
            loss, logs = m(xs)
        loss.backward()
        out = {f"x{i}": x for i, x in enumerate(xs)}
        out.update({f"noise{i}": n for i, n in enumerate(rr.rec)})
        out["loss"] = loss
        for k in ("loss_joint_recon", "loss_cross_recon", "kl_private", "kl_shared_poe", "kl_shared_uni_sum"):
            out["log." + k] = np.float32(logs[k])
        for k, p_ in m.named_parameters():
            gflat = p_.grad.reshape(-1)
            idx = grad_samples(gflat.numel())
            out["gnorm." + k] = np.float64(gflat.double().norm())
            out["gmax." + k] = np.float64(gflat.abs().max())
            out["gidx." + k] = idx
            out["gval." + k] = gflat[torch.from_numpy(idx)]
            out["wsum." + k] = np.float64(p_.detach().double().sum())     # pins the seed-regenerated weights
        mu_poe, mu_p = m.get_embedding(xs)
        out["emb_shared"] = mu_poe
        out.update({f"emb_private{i}": t for i, t in enumerate(mu_p)})
        out["meta"] = np.array([h, e, B, 3], dtype=np.int64)              # last = init seed
        out["dims"] = np.array(dims, dtype=np.int64)
        out["a"] = np.float64(a)
        save("dmvae_full_" + tag, **out)


# ----------------------------------------------------------------------------------------
# 12. DisentangledSSL, non-default branches: condzs=False, usezsx=True, distribution='normal'
#     models/disentangledssl.py:57-62,116-137, models/classifiers.py:456-459
# ----------------------------------------------------------------------------------------
def gen_dssl_variants():
    for tag, kw in {"nocond": dict(condzs=False), "zsx": dict(usezsx=True), "normal": dict(distribution="normal"),
                    "normal_nocond_zsx": dict(distribution="normal", condzs=False, usezsx=True)}.items():
        dims, h, e, B, a = [24, 40], 32, 16, 32, 0.5
        with cuda_identity_shim():
            torch.manual_seed(5)
            m = ns.disentangledssl.DisentangledSSL(output_dim=dims, hidden_dim=h, embed_dim=e, a=a, lmd_start_value=0.25, **kw)
            g = torch.Generator().manual_seed(23)
            x1, x2 = torch.randn(B, dims[0], generator=g), torch.randn(B, dims[1], generator=g)
            v1 = x1 + 0.01 * torch.randn(B, dims[0], generator=g)
            v2 = x2 + 0.01 * torch.randn(B, dims[1], generator=g)
            torch.manual_seed(1234)
            loss, logs = m(x1, x2, v1, v2)
            loss.backward()
            torch.manual_seed(1234)        # replay the reference's draw order to obtain the explicit noise
            if kw.get("distribution", "vmf") == "vmf":
                noise = [port.draw_vmf_noise(B, e, 1.0) for _ in range(4)]
            else:                          # Independent(Normal(mu, 1)).rsample(): eps = randn(mu.shape), four draws in order
                noise = [torch.randn(B, e) for _ in range(4)]
            emb_s, emb_p = m.get_embedding([x1, x2])
        out = state_arrays(m)
        out.update({f"grad.{k}": p.grad for k, p in m.named_parameters()})
        out.update(x1=x1, x2=x2, v1=v1, v2=v2, loss=loss, emb_shared=emb_s, emb_private0=emb_p[0], emb_private1=emb_p[1])
        for i, nz in enumerate(noise):
            if isinstance(nz, tuple):
                out[f"noise_w{i}"], out[f"noise_v{i}"] = nz
            else:
                out[f"noise_eps{i}"] = nz
        for k in ("shared", "clip", "loss_x", "loss_y", "specific", "ortho", "lmd"):
            out["log." + k] = np.float32(logs[k])
        out["meta"] = np.array([h, e, B], dtype=np.int64)
        out["dims"] = np.array(dims, dtype=np.int64)
        out["a"] = np.float64(a)
        out["flags"] = np.array([int(kw.get("condzs", True)), int(kw.get("usezsx", False)),
                                 int(kw.get("distribution", "vmf") == "normal")], dtype=np.int64)
        save("dssl_var_" + tag, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dssl_variants":
        gen_dssl_variants()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        gen_dmvae_full()
        sys.exit(0)
    gen_activation()
    gen_edl()
    gen_supcon()
    gen_dmvae()
    gen_dssl()
    gen_probes()
    gen_handwritten()
    gen_vmf()
    gen_eval()
    gen_datasets()
    gen_dmvae_full()
    gen_dssl_variants()
