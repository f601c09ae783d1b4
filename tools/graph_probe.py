"""Bisect which part of the training step is not CUDA-graph capturable."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import disentagled_multimodal_fusion_b200 as pkg
from disentagled_multimodal_fusion_b200 import ops
from disentagled_multimodal_fusion_b200.dp import FlatParams, GraphedStep

dev = "cuda"
torch.manual_seed(0)
B = 2048
model = pkg.DisentangledSSL(output_dim=[1024, 1024], hidden_dim=512, embed_dim=512, precision="bf16", noise_mode="device").to(dev)
probe = pkg.EvidentialProbeModule(model, num_classes=10, input_dim=512, hidden_dim=(128,), lr=3e-3, dropout=0.0,
                                  annealing_start=50, aggregation="cml", fused=1).to(dev)
probe.backbone = model
bb = FlatParams(model.parameters())
hd = FlatParams([p for n, p in probe.named_parameters() if not n.startswith("backbone.")])
x1, x2 = torch.randn(B, 1024, device=dev), torch.randn(B, 1024, device=dev)
v1, v2 = x1 + 0.01 * torch.randn_like(x1), x2 + 0.01 * torch.randn_like(x2)
y = torch.randint(0, 10, (B,), device=dev)
noise = model.draw_noise(B, dev)


def f_model():
    loss, _ = model(x1, x2, v1, v2, noise=noise)
    bb.zero_grad(); loss.backward(); bb.adam_step(1e-4, capturable=True)
    return loss.detach()


def f_probe_fwd():
    with torch.no_grad():
        l, _, _, _ = probe.shared_step([x1, x2, y])
    return l


def f_probe():
    l, _, _, _ = probe.shared_step([x1, x2, y])
    hd.zero_grad(); l.backward(); hd.adam_step(3e-3, weight_decay=1e-4, decoupled=True, capturable=True)
    return l.detach()


def f_edl():
    ev = (torch.rand(B, 3, 10, device=dev) * 3).requires_grad_()
    l, _, _ = ops.edl_fused_loss(ev, y, "cml", 5, 50, fused=1)
    l.backward()
    return l.detach()


def f_heads():
    es = torch.randn(B, 1024, device=dev)
    out = probe.x_shared(es)
    out.sum().backward()
    return out.detach().sum()


for name, fn in (("edl", f_edl), ("heads", f_heads), ("probe_fwd", f_probe_fwd), ("probe", f_probe), ("model", f_model)):
    try:
        g = GraphedStep(fn, warmup=2)
        a = float(g()); b = float(g())
        print(f"{name}: captured OK, replay -> {a:.5f} {b:.5f}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"{name}: FAILED {type(ex).__name__}: {str(ex).splitlines()[0]}", flush=True)
        tb = traceback.format_exc().splitlines()
        print("\n".join(l for l in tb if "repo/" in l or "File" in l)[-1500:], flush=True)
        break
