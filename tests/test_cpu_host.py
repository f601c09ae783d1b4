"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header declares,
the module mirrors reproduce the reference's parameter layout / init stream / noise stream, and the
data-parallel host logic works under gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.helpers import T, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    import disentagled_multimodal_fusion_b200._lib as L
    hdr = open(os.path.join(ROOT, "include", "dmf_b200.h")).read()
    names = set(re.findall(r"^\s*(?:int|long long|size_t)\s+(dmf_\w+)\s*\(", hdr, flags=re.M))
    assert len(names) >= 25
    for n in names:
        assert hasattr(L.lib, n), f"{n} declared in include/dmf_b200.h but not exported by libdmf_b200.so"
    assert set(L.EXPORTS) == names, set(L.EXPORTS) ^ names
    assert L.lib.dmf_version() == 100


def test_no_cpu_fallback():
    """Ops must fail loudly without a B200 instead of silently computing on the host."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import disentagled_multimodal_fusion_b200 as pkg
    with pytest.raises(pkg._lib.DmfError):
        pkg.ops.fuse_evidence(torch.rand(4, 2, 3), "cml")
    with pytest.raises(pkg._lib.DmfError):
        pkg.SupConLoss()(torch.rand(4, 2, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "disentagled_multimodal_fusion_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} references oracle/"


@pytest.mark.parametrize("tag", ["scene_small", "syn_small"])
def test_dmvae_init_stream_and_keys(tag):
    """Same seed => bit-equal weights and identical state_dict keys as the reference DMVAE."""
    import disentagled_multimodal_fusion_b200 as pkg
    g = load_golden("dmvae_" + tag)
    h, e, B = (int(v) for v in g["meta"])
    torch.manual_seed(3)
    m = pkg.DMVAE(output_dim=[int(d) for d in g["dims"]], a=float(g["a"]), hidden_dim=h, embed_dim=e)
    sd = m.state_dict()
    ref = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    assert set(sd) == set(ref)
    for k in sd:
        assert torch.equal(sd[k], T(ref[k])), k


def test_dssl_init_stream_keys_and_noise_stream():
    import disentagled_multimodal_fusion_b200 as pkg
    g = load_golden("dssl_small")
    h, e, B = (int(v) for v in g["meta"])
    torch.manual_seed(5)
    m = pkg.DisentangledSSL(output_dim=[int(d) for d in g["dims"]], hidden_dim=h, embed_dim=e, a=float(g["a"]),
                            lmd_start_value=0.25)
    sd = m.state_dict()
    ref = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    assert set(sd) == set(ref)
    for k in sd:
        assert torch.equal(sd[k], T(ref[k])), k
    torch.manual_seed(1234)
    noise = m.draw_noise(B, "cpu")
    for i, (w, v) in enumerate(noise):
        assert torch.equal(w, T(g[f"noise_w{i}"])) and torch.equal(v, T(g[f"noise_v{i}"]))


def test_probe_keys():
    import disentagled_multimodal_fusion_b200 as pkg
    g = load_golden("probe_cml")
    h, e, B, C = (int(v) for v in g["meta"])
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(3)
    backbone = pkg.DMVAE(output_dim=dims, a=1e-5, hidden_dim=h, embed_dim=e)
    torch.manual_seed(8)
    pm = pkg.EvidentialProbeModule(backbone, num_classes=C, input_dim=e, hidden_dim=(16,), dropout=0.1,
                                   annealing_start=50, aggregation="cml")
    ref = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    sd = pm.state_dict()
    assert set(sd) == set(ref)
    for k in sd:
        assert torch.equal(sd[k], T(ref[k])), k
    assert not any(p.requires_grad for p in pm.backbone.parameters())


def test_schedulers_and_shards():
    from disentagled_multimodal_fusion_b200.dp import shard_rows
    from disentagled_multimodal_fusion_b200.utils import ExponentialScheduler, LinearScheduler
    assert shard_rows(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        shard_rows(10, 0, 3)
    s = LinearScheduler(0.0, 1.0, 10, start_iteration=5)
    assert s(0) == 0.0 and s(20) == 1.0 and abs(s(10) - 0.5) < 1e-12
    e = ExponentialScheduler(1e-3, 1.0, 30)
    assert abs(e(15) - 10 ** -1.5) < 1e-9


_DP_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import port
from disentagled_multimodal_fusion_b200.dp import shard_rows, world
dist.init_process_group("gloo")
rank, ws = world()
torch.manual_seed(0)
B, D, T = 24, 16, 0.07
z0 = torch.nn.functional.normalize(torch.randn(B, D), dim=-1)
z1 = torch.nn.functional.normalize(torch.randn(B, D), dim=-1)
a, b = z0.clone().requires_grad_(), z1.clone().requires_grad_()
ref, _, _ = port.supcon(a, b, T)
ra, rb = torch.autograd.grad(ref, (a, b))
# sharded evaluation with exactly the decomposition ops._InfoNCE uses (row LSE per local anchor,
# gathered LSE of the other view for the transposed term, diag offset = rank * B_loc)
lo, hi = shard_rows(B, rank, ws)
l0, l1 = z0[lo:hi].contiguous(), z1[lo:hi].contiguous()
g0 = torch.empty(B, D); g1 = torch.empty(B, D)
dist.all_gather_into_tensor(g0, l0); dist.all_gather_into_tensor(g1, l1)
def rowstats(A, Bm):
    s = A @ Bm.T / T
    m = s.max(1).values
    return s, m, torch.exp(s - m[:, None]).sum(1)
s01, m01, l01 = rowstats(l0, g1); s00, m00, l00 = rowstats(l0, g0)
s10, m10, l10 = rowstats(l1, g0); s11, m11, l11 = rowstats(l1, g1)
idx = torch.arange(lo, hi)
def fin(s_c, m_c, l_c, m_i):
    mf = torch.maximum(m_c, m_i)
    sc = l_c * torch.exp(m_c - mf) + 1e-12
    pos = s_c[torch.arange(hi - lo), idx]
    return -(pos - mf - torch.log(sc)), mf + torch.log(sc)
loss0, lse0 = fin(s01, m01, l01, m00); loss1, lse1 = fin(s10, m10, l10, m11)
part = (loss0.sum() + loss1.sum()) / (2 * B)
dist.all_reduce(part)
lse0_all = torch.empty(B); lse1_all = torch.empty(B)
dist.all_gather_into_tensor(lse0_all, lse0); dist.all_gather_into_tensor(lse1_all, lse1)
coef = 1.0 / (2 * B * T)
W0 = torch.exp(s01 - lse0[:, None]) + torch.exp(s01 - lse1_all[None, :])
dz0 = coef * (W0 @ g1 - 2 * g1[lo:hi])
W1 = torch.exp(s10 - lse1[:, None]) + torch.exp(s10 - lse0_all[None, :])
dz1 = coef * (W1 @ g0 - 2 * g0[lo:hi])
assert abs(float(part) - float(ref)) < 1e-5 * abs(float(ref)), (float(part), float(ref))
assert (dz0 - ra[lo:hi]).abs().max() < 1e-5 * ra.abs().max()
assert (dz1 - rb[lo:hi]).abs().max() < 1e-5 * rb.abs().max()
# flat gradient all-reduce = SUM over ranks of per-shard grads of globally-normalised losses
w = torch.nn.Parameter(torch.ones(D))
(( (l0 * w).sum() ) / B).backward()
gsum = w.grad.clone(); dist.all_reduce(gsum)
full = torch.nn.Parameter(torch.ones(D)); ((z0 * full).sum() / B).backward()
assert torch.allclose(gsum, full.grad, atol=1e-6)
dist.destroy_process_group()
print("OK", rank)
'''


def test_data_parallel_decomposition_gloo_world2(tmp_path):
    script = tmp_path / "dp_check.py"
    script.write_text(_DP_SCRIPT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("OK") == 2
