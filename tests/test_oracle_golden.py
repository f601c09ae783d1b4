"""Pins oracle/port.py against the fixtures minted from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import port
from tests.helpers import T, assert_close, check_sampled_grads, load_golden, mlp_params_from_sd

torch.set_num_threads(1)
TOL = 2e-6


def test_activation():
    g = load_golden("activation")
    h = T(g["h"], grad=True)
    e = port.evidence_activation(h)
    assert torch.equal(e.detach(), T(g["e"]))
    (gr,) = torch.autograd.grad(e.sum(), h)
    assert_close(gr, g["grad"], TOL, "activation grad")


@pytest.mark.parametrize("tag", ["c1", "c2", "c3scene", "c3late", "c4", "ragged", "single"])
def test_edl_and_fusion(tag):
    g = load_golden("edl_" + tag)
    evid, y = T(g["evid"]), T(g["y"])
    astart = int(g["annealing_start"])
    for agg in ("cml", "avg", "joint", "disentangled", "dbf"):
        fe = port.fuse(evid, agg)
        assert torch.equal(fe, T(g["fused_" + agg])), agg
        u, ale, pred = port.uncertainty_summaries(fe)
        assert torch.equal(u, T(g["u_" + agg]))
        assert torch.equal(pred, T(g["pred_" + agg]))
        assert_close(ale, g["ale_" + agg], TOL, "aleatoric")
    for step in g["steps"]:
        for fused in (1, 0):
            ev = evid.clone().requires_grad_()
            loss = port.avg_trusted_loss(ev, y, port.fuse(ev, "cml"), fused, int(step), astart)
            (gr,) = torch.autograd.grad(loss, ev)
            assert_close(loss, g[f"loss_s{step}_f{fused}"], TOL, "loss")
            assert_close(gr, g[f"grad_s{step}_f{fused}"], TOL, "grad")


@pytest.mark.parametrize("tag", ["b64d16", "b96d64", "b33d24raw", "b2d8"])
def test_supcon_ortho(tag):
    g = load_golden("supcon_" + tag)
    z0, z1 = T(g["z0"], grad=True), T(g["z1"], grad=True)
    loss, lx, ly = port.supcon(z0, z1)
    g0, g1 = torch.autograd.grad(loss, (z0, z1))
    assert_close(loss, g["loss"], TOL, "loss")
    assert_close(lx, g["loss_x"], 1e-4, "loss_x")      # diagnostics are O(1e-3) sums of exp tails
    assert_close(ly, g["loss_y"], 1e-4, "loss_y")
    assert_close(g0, g["g0"], TOL, "g0")
    assert_close(g1, g["g1"], TOL, "g1")
    a, b = T(g["oa"], grad=True), T(g["ob"], grad=True)
    ol = port.ortho_loss(a, b)
    ga, gb = torch.autograd.grad(ol, (a, b))
    assert_close(ol, g["ortho"], TOL, "ortho")
    assert_close(ga, g["goa"], TOL, "ortho ga")
    assert_close(gb, g["gob"], TOL, "ortho gb")


def test_vmf_draw_and_householder():
    g = load_golden("vmf")
    torch.manual_seed(77)
    w, v = port.draw_vmf_noise(40, 16, 1.0)
    assert torch.equal(w, T(g["w"])) and torch.equal(v, T(g["v"]))
    e = T(g["e"], grad=True)
    z = port.vmf_rsample(e, w, v)
    assert_close(z, g["z"], TOL, "vmf z")
    (ge,) = torch.autograd.grad((z * torch.arange(16, dtype=torch.float32)).sum(), e)
    assert_close(ge, g["grad_e"], 1e-5, "vmf grad")


@pytest.mark.parametrize("tag", ["scene_small", "hw_small", "syn_small"])
def test_dmvae(tag):
    g = load_golden("dmvae_" + tag)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    N = len(g["dims"])
    enc = [mlp_params_from_sd(sd, f"encoders.{i}", (0, 2, 4), grad=True) for i in range(N)]
    dec = [mlp_params_from_sd(sd, f"decoders.{i}", (0, 2, 4), grad=True) for i in range(N)]
    xs = [T(g[f"x{i}"]) for i in range(N)]
    noise = [T(g[f"noise{i}"]) for i in range(2 * N + 1)]
    loss, logs = port.dmvae_forward(xs, enc, dec, noise, float(g["a"]))
    assert_close(loss, g["loss"], TOL, "loss")
    for k in ("loss_joint_recon", "loss_cross_recon", "kl_private", "kl_shared_poe", "kl_shared_uni_sum"):
        assert_close(logs[k], g["log." + k], TOL, k)
    loss.backward()
    for i in range(N):
        for li, idx in enumerate((0, 2, 4)):
            assert_close(enc[i][0][li].grad, g[f"grad.encoders.{i}.layers.{idx}.weight"], 1e-5, "enc wgrad")
            assert_close(dec[i][1][li].grad, g[f"grad.decoders.{i}.layers.{idx}.bias"], 1e-5, "dec bgrad")
    with torch.no_grad():
        mu, mups = port.dmvae_get_embedding(xs, enc)
    assert_close(mu, g["emb_shared"], TOL, "emb_shared")
    for i in range(N):
        assert_close(mups[i], g[f"emb_private{i}"], TOL, "emb_private")


@pytest.mark.parametrize("tag", ["c1_hw", "c3_cub", "c2_syn"])
def test_dmvae_full_size(tag):
    """DMVAE at the BASELINE.json sizes (C1 HandWritten rows h=512 e=200 B=100; C3 CUB 1024/300; C2 synthetic B=4096).
    The weights are regenerated from the init seed (the mirror's init stream is seed-equal to the reference's) and
    pinned through their per-tensor sums."""
    import disentagled_multimodal_fusion_b200 as pkg
    g = load_golden("dmvae_full_" + tag)
    h, e, B, seed = (int(v) for v in g["meta"])
    dims = [int(d) for d in g["dims"]]
    N = len(dims)
    torch.manual_seed(seed)
    m = pkg.DMVAE(output_dim=dims, a=float(g["a"]), hidden_dim=h, embed_dim=e)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in m.named_parameters():
        assert abs(float(v.detach().double().sum()) - float(g["wsum." + k])) < 1e-9 * max(1.0, abs(float(g["wsum." + k]))) + 1e-9, k
    torch.set_num_threads(max(1, min(8, torch.get_num_threads() if torch.get_num_threads() > 1 else 8)))
    try:
        enc = [mlp_params_from_sd(sd, f"encoders.{i}", (0, 2, 4), grad=True) for i in range(N)]
        dec = [mlp_params_from_sd(sd, f"decoders.{i}", (0, 2, 4), grad=True) for i in range(N)]
        xs = [T(g[f"x{i}"]) for i in range(N)]
        noise = [T(g[f"noise{i}"]) for i in range(2 * N + 1)]
        loss, logs = port.dmvae_forward(xs, enc, dec, noise, float(g["a"]))
        assert_close(loss, g["loss"], 1e-5, "loss")
        for k in ("loss_joint_recon", "loss_cross_recon", "kl_private", "kl_shared_poe", "kl_shared_uni_sum"):
            assert_close(logs[k], g["log." + k], 1e-5, k)
        loss.backward()
        grads = {}
        for i in range(N):
            for li, idx in enumerate((0, 2, 4)):
                grads[f"encoders.{i}.layers.{idx}.weight"] = enc[i][0][li].grad
                grads[f"encoders.{i}.layers.{idx}.bias"] = enc[i][1][li].grad
                grads[f"decoders.{i}.layers.{idx}.weight"] = dec[i][0][li].grad
                grads[f"decoders.{i}.layers.{idx}.bias"] = dec[i][1][li].grad
        check_sampled_grads(grads, g, 2e-5)
        with torch.no_grad():
            mu, mups = port.dmvae_get_embedding(xs, enc)
        assert_close(mu, g["emb_shared"], 1e-5, "emb_shared")
        for i in range(N):
            assert_close(mups[i], g[f"emb_private{i}"], 1e-5, "emb_private")
    finally:
        torch.set_num_threads(1)


def _dssl_noise(g):
    if "noise_eps0" in g:
        return [T(g[f"noise_eps{i}"]) for i in range(4)]
    return [(T(g[f"noise_w{i}"]), T(g[f"noise_v{i}"])) for i in range(4)]


@pytest.mark.parametrize("tag", ["nocond", "zsx", "normal", "normal_nocond_zsx"])
def test_dssl_variants(tag):
    """non-default branches of DisentangledSSL (condzs=False, usezsx=True, distribution='normal')"""
    g = load_golden("dssl_var_" + tag)
    condzs, usezsx, normal = (bool(v) for v in g["flags"])
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    names = {"x1s": "encoder_x1s", "x2s": "encoder_x2s", "x1": "encoder_x1", "x2": "encoder_x2"}
    p = {k: mlp_params_from_sd(sd, n, (0, 2, 4), grad=True) for k, n in names.items()}
    x1, x2, v1, v2 = (T(g[k]) for k in ("x1", "x2", "v1", "v2"))
    loss, logs = port.dssl_forward(x1, x2, v1, v2, p, _dssl_noise(g), a=float(g["a"]), lmd=float(g["log.lmd"]), condzs=condzs,
                                   usezsx=usezsx, distribution="normal" if normal else "vmf")
    assert_close(loss, g["loss"], TOL, "loss")
    for k in ("shared", "specific", "ortho"):
        assert_close(logs[k], g["log." + k], TOL, k)
    loss.backward()
    for k, n in names.items():
        for li, idx in enumerate((0, 2, 4)):
            assert_close(p[k][0][li].grad, g[f"grad.{n}.layers.{idx}.weight"], 2e-5, f"{n} wgrad {idx}")
    with torch.no_grad():
        es, ep = port.dssl_get_embedding(x1, x2, p, condzs=condzs)
    assert_close(es, g["emb_shared"], TOL, "emb_shared")
    assert_close(ep[0], g["emb_private0"], TOL, "emb_private0")


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_dssl(tag):
    g = load_golden("dssl_" + tag)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    names = {"x1s": "encoder_x1s", "x2s": "encoder_x2s", "x1": "encoder_x1", "x2": "encoder_x2"}
    p = {k: mlp_params_from_sd(sd, n, (0, 2, 4), grad=True) for k, n in names.items()}
    noise = [(T(g[f"noise_w{i}"]), T(g[f"noise_v{i}"])) for i in range(4)]
    x1, x2, v1, v2 = (T(g[k]) for k in ("x1", "x2", "v1", "v2"))
    loss, logs = port.dssl_forward(x1, x2, v1, v2, p, noise, a=float(g["a"]), lmd=float(g["log.lmd"]))
    assert_close(loss, g["loss"], TOL, "loss")
    for k in ("shared", "specific", "ortho"):
        assert_close(logs[k], g["log." + k], TOL, k)
    assert_close(logs["loss_x"], g["log.loss_x"], 1e-4, "loss_x")
    loss.backward()
    for k, n in names.items():
        for li, idx in enumerate((0, 2, 4)):
            assert_close(p[k][0][li].grad, g[f"grad.{n}.layers.{idx}.weight"], 2e-5, f"{n} wgrad {idx}")
            assert_close(p[k][1][li].grad, g[f"grad.{n}.layers.{idx}.bias"], 2e-5, f"{n} bgrad {idx}")
    with torch.no_grad():
        es, ep = port.dssl_get_embedding(x1, x2, p)
    assert_close(es, g["emb_shared"], TOL, "emb_shared")
    assert_close(ep[1], g["emb_private1"], TOL, "emb_private1")


@pytest.mark.parametrize("name,heads_prefix,agg,uses_shared", [
    ("probe_cml", None, "cml", True), ("probe_avg", None, "avg", True),
    ("probe_joint", None, "joint", True), ("probe_disentangled", None, "disentangled", True),
    ("probe_dis_cml", "spec_heads", "cml", False),
])
def test_probes(name, heads_prefix, agg, uses_shared):
    g = load_golden(name)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    N = len(g["dims"])
    enc = [mlp_params_from_sd(sd, f"backbone.encoders.{i}", (0, 2, 4)) for i in range(N)]
    xs = [T(g[f"x{i}"]) for i in range(N)]
    with torch.no_grad():
        mu, mups = port.dmvae_get_embedding(xs, enc)
    if uses_shared:
        heads = [mlp_params_from_sd(sd, "x_shared", (0, 3), grad=True)] + \
                [mlp_params_from_sd(sd, f"x_specs.{i}", (0, 3), grad=True) for i in range(N)]
        embeds = [mu] + mups
        fused = 1
    else:
        heads = [mlp_params_from_sd(sd, f"spec_heads.{i}", (0, 3), grad=True) for i in range(N)]
        embeds = mups
        fused = 1
    loss, ea, _, ev = port.probe_shared_step(embeds, heads, T(g["y"]), agg, fused,
                                             int(g["annealing_step"]), int(g["annealing_start"]))
    assert_close(loss, g["loss"], TOL, "loss")
    assert_close(ev, g["evidences"], TOL, "evidences")
    assert_close(ea, g["evidences_a"], TOL, "evidences_a")
    assert torch.equal(ea.argmax(-1), T(g["evidences_a"]).argmax(-1))
    loss.backward()
    gname = "x_shared.layers.0.weight" if uses_shared else "spec_heads.0.layers.0.weight"
    assert_close(heads[0][0][0].grad, g["grad." + gname], 1e-5, "head wgrad")


@pytest.mark.parametrize("name", ["latefusion_dbf", "latefusion_cml", "latefusion_avg", "latefusion_handwritten"])
def test_latefusion(name):
    g = load_golden(name)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    N = len(g["dims"])
    agg = name.split("_")[1] if "handwritten" not in name else "cml"
    heads = [mlp_params_from_sd(sd, f"heads.{i}", (0, 3), grad=True) for i in range(N)]
    xs = [T(g[f"x{i}"]) for i in range(N)]
    loss, ea, _, ev = port.probe_shared_step(xs, heads, T(g["y"]), agg, 1,
                                             int(g["annealing_step"]), int(g["annealing_start"]))
    assert_close(loss, g["loss"], TOL, "loss")
    assert_close(ev, g["evidences"], TOL, "evidences")
    assert_close(ea, g["evidences_a"], 5e-6, "evidences_a")
    loss.backward()
    assert_close(heads[0][0][0].grad, g["grad.heads.0.layers.0.weight"], 1e-5, "head wgrad")


def _reduce_dict(acc, N, with_shared):
    """Build the reference's result dict from the oracle accumulators (reduce_block of analysis.py:157-169)."""
    def block(st):
        c, ev, epi, ale, n_inc, iev, iepi, iale = [float(x) for x in st]
        return {"accuracy": c / N, "evidence_mean": ev / N, "epistemic_mean": epi / N, "aleatoric_mean": ale / N,
                "incorrect_only": {"evidence_mean": iev / n_inc if n_inc > 0 else 0.0,
                                   "epistemic_mean": iepi / n_inc if n_inc > 0 else 0.0,
                                   "aleatoric_mean": iale / n_inc if n_inc > 0 else 0.0}}
    V = acc["stats"].shape[0] - 1
    unc = acc["class_sum"] / N
    tru = acc["true_sum"] / torch.clamp(acc["class_counts"], min=1e-12)
    lo = 1 if with_shared else 0
    out = {"per_view": [block(acc["stats"][v]) for v in range(lo, V)], "fused": block(acc["stats"][V]),
           "per_class_evidence": {"unconditional": {"per_view": [unc[v].tolist() for v in range(lo, V)], "fused": unc[V].tolist()},
                                  "true_class": {"per_view": [tru[v].tolist() for v in range(lo, V)], "fused": tru[V].tolist()}}}
    if with_shared:
        out["shared"] = block(acc["stats"][0])
        out["per_class_evidence"]["unconditional"]["shared"] = unc[0].tolist()
        out["per_class_evidence"]["true_class"]["shared"] = tru[0].tolist()
    return out


def _flat(d, prefix=""):
    out = {}
    if isinstance(d, dict):
        for k, v in d.items():
            out.update(_flat(v, prefix + str(k) + "."))
    elif isinstance(d, (list, tuple)):
        for i, v in enumerate(d):
            out.update(_flat(v, prefix + str(i) + "."))
    else:
        out[prefix] = float(d)
    return out


def assert_result_dicts_close(ours, ref, rtol=1e-5):
    fo, fr = _flat(ours), _flat(ref)
    assert fo.keys() == fr.keys(), set(fo) ^ set(fr)
    for k in fr:
        assert abs(fo[k] - fr[k]) <= rtol * max(1.0, abs(fr[k])), (k, fo[k], fr[k])


def test_eval_reducer_against_reference_dicts():
    """oracle.port.eval_reduce accumulated over the ragged batches == the dicts the UNMODIFIED reference functions
    (analysis.py:5-399) returned for the same evidences (fixture eval_reduce.npz)."""
    import json
    g = load_golden("eval_reduce")
    evid, y, fused = T(g["evid"]), T(g["y"]), T(g["fused"])
    acc, o = None, 0
    for b in g["sizes"]:
        b = int(b)
        part = port.eval_reduce(evid[o:o + b], fused[o:o + b], y[o:o + b])
        acc = part if acc is None else {k: acc[k] + part[k] for k in acc}
        o += b
    N = evid.shape[0]
    assert_result_dicts_close(_reduce_dict(acc, N, False), json.loads(bytes(g["plain_json"]).decode()))
    assert_result_dicts_close(_reduce_dict(acc, N, True), json.loads(bytes(g["shared_json"]).decode()))
