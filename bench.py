#!/usr/bin/env python
"""bench.py -- train samples/sec of the encoder + InfoNCE + EDL step (BASELINE.json metric).

Workload (config C5 of SURVEY §8): DisentangledSSL 2-view, 1024-d per view, hidden 512, embed 512,
T=0.07, global batch 65536 sharded by rows over the ranks (data parallel: embeddings + row-LSEs
all-gathered with NCCL inside the InfoNCE op, ortho Gram and flat gradients all-reduced), followed by
an evidential probe (3 heads, C=10, hidden 128, cml fusion, fused EDL kernel) on the same batch.
One step = backbone fwd + bwd + fused Adam, then probe fwd + bwd + fused AdamW.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # CPU oracle port of the reference path

Prints ONE JSON line (rank 0).  `value` = device-resident inputs, CUDA-event timed, max over ranks;
`e2e` = same step fed from pinned host buffers (H2D inside the timed region) with a device->host read
of the loss every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "train samples/sec (encoder+InfoNCE+EDL step)"
UNIT = "samples/s"
DIMS, HID, EMB, NCLS, TEMP = [1024, 1024], 512, 512, 10, 0.07


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): ONE background
    `nvidia-smi -lms 200` process whose output is parsed at the end (no periodic fork from the timing process)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.samples, self.reasons, self.sm_max = [], set(), None
        self.stop_flag = False

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
        except Exception:  # noqa: BLE001
            self.proc = None

    def join(self, timeout=None):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=timeout or 5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.sm_max = float(f[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(self.NAMES, f[2:]):
                if v.lower() == "active":
                    self.reasons.add(n)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step, timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(B, seed=0):
    from oracle import port
    g = torch.Generator().manual_seed(seed)
    names = {"x1s": (DIMS[0], EMB), "x2s": (DIMS[1], EMB), "x1": (DIMS[0] + EMB, EMB), "x2": (DIMS[1] + EMB, EMB)}
    p = {k: port.xavier_mlp_params((din, HID, HID), dout, g) for k, (din, dout) in names.items()}
    heads = [port.xavier_mlp_params((2 * EMB, 128), NCLS, g)] + [port.xavier_mlp_params((EMB, 128), NCLS, g) for _ in range(2)]
    bb_params = [t for ws, bs in p.values() for t in ws + bs]
    hd_params = [t for ws, bs in heads for t in ws + bs]
    opt_b = torch.optim.Adam(bb_params, lr=1e-4)
    opt_h = torch.optim.AdamW(hd_params, lr=3e-3, weight_decay=1e-4)
    x1, x2 = torch.randn(B, DIMS[0], generator=g), torch.randn(B, DIMS[1], generator=g)
    v1, v2 = x1 + 0.01 * torch.randn(B, DIMS[0], generator=g), x2 + 0.01 * torch.randn(B, DIMS[1], generator=g)
    y = torch.randint(0, NCLS, (B,), generator=g)

    def step():
        noise = [port.draw_vmf_noise(B, EMB, 1.0) for _ in range(4)]
        loss, _ = port.dssl_forward(x1, x2, v1, v2, p, noise, a=1.0, lmd=0.0, T=TEMP)
        opt_b.zero_grad(set_to_none=True)
        loss.backward()
        opt_b.step()
        with torch.no_grad():
            es, ep = port.dssl_get_embedding(x1, x2, p)
        l2, _, _, _ = port.probe_shared_step([es] + ep, heads, y, "cml", 1, 5, 50)
        opt_h.zero_grad(set_to_none=True)
        l2.backward()
        opt_h.step()
        return float(loss) + float(l2)
    return step


def time_cpu(B, steps, warmup):
    step = cpu_step_factory(B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt


def host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1: give the CPU arm every core this process may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_float32_matmul_precision("highest")
    B = args.cpu_batch
    cores = host_threads()
    val, dt = time_cpu(B, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 DisentangledSSL 2x1024-d, hidden 512, embed 512, T=0.07 + 3-head evidential probe",
                   "global_batch": B, "note": "CPU oracle port of the reference step (oracle/port.py, torch CPU fp32 'highest'); "
                   "the reference materialises [2B,2B] logits so it is timed at a bounded batch; cost grows ~B^2"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full steps (fwd+bwd+Adam, probe fwd+bwd+AdamW) at batch {B}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _timeit(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def kernel_rooflines(dev, Bl, pk):
    """Live rooflines of K1 (grouped MLP GEMM, tensor bound) and K3 (evidence fusion + EDL, HBM bound), each kernel
    timed alone with CUDA events on the current stream; inputs rotate through buffers larger than the L2."""
    from disentagled_multimodal_fusion_b200 import ops, _lib as L
    from disentagled_multimodal_fusion_b200._lib import lib, check, ptr, stream
    M = 2 * Bl
    per_shape, flops_tot, ms_tot = {}, 0.0, 0.0
    for K in (1024, 512, 1536):
        A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
        W = [torch.randn(HID, K, device=dev).bfloat16() for _ in range(2)]
        o = [torch.empty(M, HID, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        bias = torch.zeros(HID, device=dev)
        descs = [dict(A=A[g], lda=K, B=W[g], ldb=K, out_bf16=o[g], ldo_bf16=HID, bias=bias, M=M, N=HID, K=K) for g in range(2)]
        ms = _timeit(lambda: ops.gemm_tc(descs, L.EPI_BIAS_RELU), 10, 3)
        fl = 2.0 * 2 * M * HID * K
        per_shape[f"K={K}"] = round(fl / ms / 1e9, 1)
        flops_tot += fl
        ms_tot += ms
        del A, W, o
    k1 = {"kernel": "gemm_bf16_tc2_kernel", "bound": "tensor", "achieved": flops_tot / ms_tot / 1e9, "peak": pk["tf_burst"],
          "unit": "TFLOP/s", "frac": flops_tot / ms_tot / 1e9 / pk["tf_burst"], "traffic": None,
          "peak_source": pk["src"] + " burst bf16 (kernel timed alone)",
          "shape": f"2 groups, M={M}, N={HID}, K in (1024, 512, 1536), bias+ReLU epilogue, bf16 out; FLOP-weighted over the three layers",
          "tflops_by_layer": per_shape}
    Be, V, C = 1 << 22, 4, 42
    evid = torch.rand(Be, V, C, device=dev) * 3
    y = torch.randint(0, C, (Be,), device=dev)
    fused = torch.empty(Be, C, device=dev)
    grad = torch.empty_like(evid)
    parts = torch.zeros(4, device=dev)
    p = L.EdlParams(Be, V, C, 0, 0.5, 0.6, 1.0 / Be)
    ms = _timeit(lambda: check(lib.dmf_edl_fused(ptr(evid), ptr(y), p, 0, ptr(fused), ptr(grad), 0, 0, 0, ptr(parts), stream())), 5, 2)
    byt = Be * (8.0 * V * C + 4 * C + 16)
    cap = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.isfile(tpath):
        try:
            cap = json.load(open(tpath)).get("edl_fused_kernel", {})
        except Exception:  # noqa: BLE001
            cap = {}
    k3 = {"kernel": "edl_fused_kernel (training mode: gradient + loss + conflict term)", "bound": "hbm",
          "achieved": byt / ms / 1e6, "peak": pk["hbm"], "unit": "GB/s", "frac": byt / ms / 1e6 / pk["hbm"],
          "traffic": cap.get("dram_bytes"), "avg_ms": ms, "peak_source": pk["src"] + " HBM copy bandwidth",
          "shape": f"B={Be}, V={V}, C={C}; algorithmic bytes/sample = 8VC + 4C + 16 = {8 * V * C + 4 * C + 16}"}
    del evid, grad, fused
    torch.cuda.empty_cache()
    return k1, k3


def gpu_eager_context(dev, B):
    """The oracle port of the reference step executed by torch-eager ATen kernels on the GPU (context only: this is
    what the unmodified reference would run on this device; it is neither the product nor the stated baseline)."""
    from oracle import port
    g = torch.Generator().manual_seed(0)
    names = {"x1s": (DIMS[0], EMB), "x2s": (DIMS[1], EMB), "x1": (DIMS[0] + EMB, EMB), "x2": (DIMS[1] + EMB, EMB)}
    p = {k: port.xavier_mlp_params((din, HID, HID), dout, g) for k, (din, dout) in names.items()}
    p = {k: ([w.detach().to(dev).requires_grad_() for w in ws], [b.detach().to(dev).requires_grad_() for b in bs]) for k, (ws, bs) in p.items()}
    params = [t for ws, bs in p.values() for t in ws + bs]
    opt = torch.optim.Adam(params, lr=1e-4)
    x1, x2 = torch.randn(B, DIMS[0], generator=g).to(dev), torch.randn(B, DIMS[1], generator=g).to(dev)
    v1, v2 = x1 + 0.01 * torch.randn_like(x1), x2 + 0.01 * torch.randn_like(x2)
    noise = [tuple(t.to(dev) for t in port.draw_vmf_noise(B, EMB, 1.0)) for _ in range(4)]

    def step():
        loss, _ = port.dssl_forward(x1, x2, v1, v2, p, noise, a=1.0, lmd=0.0, T=TEMP)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    ms = _timeit(step, 3, 2)
    return {"value": B / (ms / 1e3), "unit": UNIT, "batch": B, "ms_per_step": ms,
            "what": "oracle port of the DSSL step (fp32, torch-eager ATen/cuBLAS kernels, vMF noise pre-drawn, no probe) on this GPU"}


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs C1-C4 (small, launch-bound shapes): samples/s + launches/step beside the CPU port at the SAME batch
# ------------------------------------------------------------------------------------------------
SMALL = {
    "c1": dict(workload="C1 HandWritten-shape DMVAE (6 views, h=512, e=200, a=1e-5, B=100) + evidential probe (V=7, C=10, hidden 128, cml)",
               dims=[240, 76, 216, 47, 64, 6], e=200, B=100, C=10, head="probe", agg="cml", fused=1),
    "c2": dict(workload="C2 run_synthetic DMVAE (2 views x 32-d, h=512, e=16, B=4096) + evidential probe (V=3, C=3, fused=0)",
               dims=[32, 32], e=16, B=4096, C=3, head="probe", agg="cml", fused=0),
    "c3": dict(workload="C3 CUB-shape (GoogLeNet 1024-d + doc2vec 300-d, C=10, B=100): DMVAE step + LateFusion with "
                        "discounted-belief (Dempster-Shafer conflict) fusion",
               dims=[1024, 300], e=200, B=100, C=10, head="latefusion", agg="dbf", fused=1),
    "c4": dict(workload="C4 LUMA-shape evidential fusion only: evidences [B=64, V=4, C=42] -> fused evidence + AvgTrustedLoss + gradient",
               dims=None, e=0, B=64, C=42, V=4, head="edl", agg="cml", fused=1),
}


def run_small_config(args):
    import disentagled_multimodal_fusion_b200 as pkg
    from disentagled_multimodal_fusion_b200 import ops, _lib
    from disentagled_multimodal_fusion_b200.dp import FlatParams
    from oracle import port        # CPU baseline leg only (timed beside, never on the product path)
    cfg = SMALL[args.config]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.require_device()
    B, C = cfg["B"], cfg["C"]
    g = torch.Generator().manual_seed(1234)
    y = torch.randint(0, C, (B,), generator=g)
    K, W = max(args.steps, 20), max(args.warmup, 3)

    if cfg["head"] == "edl":
        V = cfg["V"]
        evid_h = torch.exp(torch.clamp(torch.randn(B, V, C, generator=g) * 2, -10, 10))
        evid = evid_h.to(dev).requires_grad_()
        yd = y.to(dev)

        def gstep():
            evid.grad = None
            loss, fe, _ = ops.edl_fused_loss(evid, yd, cfg["agg"], 5, 50, fused=cfg["fused"])
            loss.backward()
            return loss
        ev_c = evid_h.clone().requires_grad_()

        def cstep():
            ev_c.grad = None
            loss = port.avg_trusted_loss(ev_c, y, port.fuse(ev_c, cfg["agg"]), cfg["fused"], 5, 50)
            loss.backward()
            return loss
        h2d = evid_h.numel() * 4 + B * 8
        host_in, dev_in = [evid_h.pin_memory(), y.pin_memory()], None
    else:
        dims, e = cfg["dims"], cfg["e"]
        xs_h = [torch.rand(B, d, generator=g) for d in dims]
        torch.manual_seed(0)
        model = pkg.DMVAE(output_dim=dims, a=1e-5, hidden_dim=512, embed_dim=e).to(dev)
        if cfg["head"] == "probe":
            head = pkg.EvidentialProbeModule(model, num_classes=C, input_dim=e, hidden_dim=(128,), dropout=0.0,
                                             annealing_start=50, aggregation=cfg["agg"], fused=cfg["fused"]).to(dev)
            head.backbone = model
            hparams = [p for n, p in head.named_parameters() if not n.startswith("backbone.")]
        else:
            head = pkg.LateFusion([(pkg.IdentityEncoder, {}) for _ in dims], dims, C, dropout=0.0, aggregation=cfg["agg"],
                                  annealing_start=50, hidden_dim=(128,), fused=cfg["fused"]).to(dev)
            hparams = list(head.parameters())
        head.criterion.annealing_step = 5
        bb, hd = FlatParams(model.parameters()), FlatParams(hparams)
        xs, yd = [x.to(dev) for x in xs_h], y.to(dev)

        def gstep(inp=None, capturable=False):
            xin = xs if inp is None else inp[:-1]
            yin = yd if inp is None else inp[-1]
            loss, _ = model(xin)
            bb.zero_grad()
            loss.backward()
            bb.adam_step(1e-4, capturable=capturable)
            pl = head.shared_step([*xin, yin])[0]
            hd.zero_grad()
            pl.backward()
            hd.adam_step(3e-3, weight_decay=1e-4, decoupled=True, capturable=capturable)
            return loss.detach() + pl.detach()
        # CPU port of the same step at the same batch
        gp = torch.Generator().manual_seed(0)
        enc = [port.xavier_mlp_params((d, 512, 512), 4 * e, gp) for d in dims]
        dec = [port.xavier_mlp_params((2 * e, 512, 512), d, gp) for d in dims]
        if cfg["head"] == "probe":
            heads = [port.xavier_mlp_params((e, 128), C, gp) for _ in range(len(dims) + 1)]
        else:
            heads = [port.xavier_mlp_params((d, 128), C, gp) for d in dims]
        opt_b = torch.optim.Adam([t for ws, bs in enc + dec for t in ws + bs], lr=1e-4)
        opt_h = torch.optim.AdamW([t for ws, bs in heads for t in ws + bs], lr=3e-3, weight_decay=1e-4)

        def cstep():
            noise = [torch.randn(B, e) for _ in range(2 * len(dims) + 1)]
            loss, _ = port.dmvae_forward(xs_h, enc, dec, noise, 1e-5)
            opt_b.zero_grad(set_to_none=True)
            loss.backward()
            opt_b.step()
            if cfg["head"] == "probe":
                with torch.no_grad():
                    mu, mups = port.dmvae_get_embedding(xs_h, enc)
                emb = [mu] + mups
            else:
                emb = xs_h
            l2 = port.probe_shared_step(emb, heads, y, cfg["agg"], cfg["fused"], 5, 50)[0]
            opt_h.zero_grad(set_to_none=True)
            l2.backward()
            opt_h.step()
            return loss
        h2d = sum(x.numel() * 4 for x in xs_h) + B * 8
        host_in = [x.pin_memory() for x in xs_h] + [y.pin_memory()]
        dev_in = [x.to(dev) for x in host_in]

    for _ in range(W):
        gstep()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    gstep()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    ms_eager = _timeit(gstep, K, 0)
    # the same step as ONE CUDA graph (forward, backward, fused optimizers; torch's graph-safe Philox state feeds the
    # reparameterisation noise): these shapes are launch-bound, the graph removes the host from the loop
    launch_mode, replay = "eager", gstep
    if args.graph != "off" and cfg["head"] != "edl":
        try:
            from disentagled_multimodal_fusion_b200.dp import GraphedStep
            graphed = GraphedStep(lambda: gstep(dev_in, capturable=True), warmup=2, stream=torch.cuda.current_stream())
            torch.cuda.synchronize()
            replay, launch_mode = graphed, "cuda_graph"
        except Exception as ex:  # noqa: BLE001
            if args.graph == "on":
                raise
            launch_mode = f"eager (graph capture failed: {type(ex).__name__})"
            torch.cuda.synchronize()
    for _ in range(3):
        replay()
    sampler = ClockSampler(0)
    sampler.start()
    ms = _timeit(replay, K, 0)
    sampler.join(timeout=3)

    # end to end: pinned host batch -> device every step, loss read back on the host every step
    def e2e_step():
        if dev_in is not None:
            for d, h in zip(dev_in, host_in):
                d.copy_(h, non_blocking=True)
            out = replay() if launch_mode == "cuda_graph" else gstep(dev_in)
        else:
            evid.data.copy_(host_in[0], non_blocking=True)
            out = gstep()
        return float(out)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / K

    cores = host_threads()
    torch.set_float32_matmul_precision("highest")
    for _ in range(2):
        cstep()
    t0 = time.perf_counter()
    nc = 0
    while nc < 5 or (time.perf_counter() - t0 < 5.0 and nc < 200):
        cstep()
        nc += 1
    cdt = (time.perf_counter() - t0) / nc
    line = {
        "metric": METRIC, "value": B / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "global_batch": B, "launch": launch_mode, "note": "launch-bound shape: "
                   "report samples/s and launches/step, not a roofline fraction (SURVEY 8d)"},
        "eager_ms_per_step": ms_eager,
        "clocks": sampler.summary(),
        "e2e": {"value": B / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches) * K, "gpu_launches_per_step": int(launches),
        "roofline": None,
        "cpu_baseline": {"value": B / cdt, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": cdt * 1e3,
                         "sample": f"{nc} full steps of the oracle port at the same batch {B}"},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_local_cpus(dev_index):
    """CPUs NVML reports as local to this GPU (its NUMA node), restricted to the CPUs this process may use."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = getattr(torch.cuda.get_device_properties(dev_index), "uuid", None)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode()) if uuid is not None else None
        except Exception:  # noqa: BLE001
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(dev_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


def run_gpu(args, rank, world, local_rank):
    import disentagled_multimodal_fusion_b200 as pkg
    from disentagled_multimodal_fusion_b200 import ops, _lib
    from disentagled_multimodal_fusion_b200.dp import FlatParams, GraphedStep, shard_rows

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.require_device()
    Bg = args.batch
    lo, hi = shard_rows(Bg, rank, world)
    Bl = hi - lo
    prec = args.precision

    # K1 / K3 rooflines first: each kernel timed ALONE on the still-cool GPU (their denominators are the burst peaks)
    roof_k1 = roof_k3 = None
    if world == 1 and not args.no_kernel_rooflines:
        roof_k1, roof_k3 = kernel_rooflines(dev, Bl, peaks())

    torch.manual_seed(0)                      # identical replicated parameters on every rank
    model = pkg.DisentangledSSL(output_dim=DIMS, hidden_dim=HID, embed_dim=EMB, a=1.0, vmfkappa=1, precision=prec,
                                noise_mode="device").to(dev)
    probe = pkg.EvidentialProbeModule(model, num_classes=NCLS, input_dim=EMB, hidden_dim=(128,), lr=3e-3, dropout=0.0,
                                      annealing_start=50, aggregation="cml", fused=1, precision=prec).to(dev)
    probe.backbone = model                    # the probe reads the live (frozen-for-it) backbone of this step
    probe.criterion.annealing_step = 5
    bb = FlatParams(model.parameters())
    hd = FlatParams([p for n, p in probe.named_parameters() if not n.startswith("backbone.")])

    # Rank-independent synthetic data: the GLOBAL batch is cut into 8 fixed chunks, chunk c drawn from seed 1234 + c,
    # so that N = 1, 2, 4, 8 all train on the same global batch (loss_check below compares them).
    def host_chunk(c, rows):
        g = torch.Generator().manual_seed(1234 + c)
        return {"x1": torch.randn(rows, DIMS[0], generator=g), "x2": torch.randn(rows, DIMS[1], generator=g),
                "y": torch.randint(0, NCLS, (rows,), generator=g)}

    def host_rows(r, w):
        if Bg % 8 == 0 and 8 % w == 0:
            parts = [host_chunk(c, Bg // 8) for c in range(r * 8 // w, (r + 1) * 8 // w)]
            return {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
        return host_chunk(100 + r, Bg // w)
    # One process per GPU: the pinned staging buffers are allocated while the process is bound to the CPUs of the GPU's
    # own NUMA node (first touch), so the H2D DMA of every rank reads node-local memory instead of crossing the socket
    # interconnect.  The binding is dropped again before the CPU baseline arm (which uses every host thread).
    rows_h = host_rows(rank, world)
    host_unbound = {k: v.pin_memory() for k, v in rows_h.items()} if args.e2e_diag else None
    all_cpus = os.sched_getaffinity(0)
    local_cpus = gpu_local_cpus(local_rank) if (world > 1 and not args.no_numa_bind) else None
    if local_cpus:
        os.sched_setaffinity(0, local_cpus)
    host = {k: v.clone().pin_memory() for k, v in rows_h.items()}
    del rows_h
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # The batch a user hands to training_step is (x1, x2[, y]); the augmented views v1, v2 are made ON THE DEVICE by
    # dmf_augment (DisentangledSSL.shared_step, noise_mode="device": SURVEY 8f-1 -- the reference's augment_data is an
    # O(B) host loop).  Like the vMF noise, the two augmentation launches run before every step, outside the graph
    # (host-side seed counter), into fixed buffers.
    aug_seed = [0]

    def with_views(bufset):
        bufset["v1"] = torch.empty_like(bufset["x1"])
        bufset["v2"] = torch.empty_like(bufset["x2"])
        return bufset

    def augment_views(bufset):
        aug_seed[0] += 1
        ops.augment(bufset["x1"], 0xA06 + aug_seed[0], 0, out=bufset["v1"])
        ops.augment(bufset["x2"], 0xA06 + aug_seed[0], 1, out=bufset["v2"])
    devin = with_views({k: v.to(dev) for k, v in host.items()})
    augment_views(devin)

    noise_bufs = model.draw_noise(Bl, dev)          # fixed buffers, refilled in place every step (outside the graph)

    def step(inp, capturable=False):
        loss, logs = model(inp["x1"], inp["x2"], inp["v1"], inp["v2"], noise=noise_bufs)
        bb.zero_grad()
        loss.backward()
        # the backbone's gradient all-reduce runs UNDER the probe's forward / backward (the probe reads the backbone
        # this step's forward pass used; both optimizers step at the end)
        pending = bb.allreduce_grads(async_op=True)
        ploss, _, _, _ = probe.shared_step([inp["x1"], inp["x2"], inp["y"]])
        hd.zero_grad()
        ploss.backward()
        hd.allreduce_grads()
        bb.wait_grads(pending)
        bb.adam_step(1e-4, capturable=capturable)
        hd.adam_step(3e-3, weight_decay=1e-4, decoupled=True, capturable=capturable)
        return loss.detach() + ploss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- loss_check: ONE step without optimizer update on the fresh (seed 0) parameters, rank-independent data AND
    #      rank-independent noise (noise / augmentation drawn for the GLOBAL batch on every rank, local rows sliced):
    #      global loss, probe loss and the norms of the all-reduced gradients must agree between N = 1, 2, 4, 8
    def loss_check():
        full = host_rows(0, 1) if world > 1 else host
        gx1, gx2 = full["x1"].to(dev), full["x2"].to(dev)
        gv1 = ops.augment(gx1, 0xC0FFEE, 0)
        gv2 = ops.augment(gx2, 0xC0FFEE, 1)
        seed_keep = model.noise_seed
        model.noise_seed = 0x777
        gnoise = model.draw_noise(Bg, dev)
        model.noise_seed = seed_keep
        inp = {"x1": gx1[lo:hi], "x2": gx2[lo:hi], "v1": gv1[lo:hi].contiguous(), "v2": gv2[lo:hi].contiguous(),
               "y": full["y"].to(dev)[lo:hi]}
        nz = [(w[lo:hi].contiguous(), v[lo:hi].contiguous()) for w, v in gnoise]
        loss, logs = model(inp["x1"], inp["x2"], inp["v1"], inp["v2"], noise=nz)
        bb.zero_grad()
        loss.backward()
        bb.allreduce_grads()
        ploss = probe.shared_step([inp["x1"], inp["x2"], inp["y"]])[0]
        hd.zero_grad()
        ploss.backward()
        hd.allreduce_grads()
        out = {"loss": float(loss.detach()), "shared": float(logs["shared"]), "specific": float(logs["specific"]),
               "probe_loss": float(ploss.detach()), "grad_norm": float(bb.grad.norm()), "probe_grad_norm": float(hd.grad.norm()),
               "what": "one step on the fresh parameters; data, augmentation and vMF noise identical for every N"}
        bb.zero_grad()
        hd.zero_grad()
        del gx1, gx2, gv1, gv2, gnoise
        torch.cuda.empty_cache()
        return out
    lcheck = loss_check() if args.loss_check else None

    step_marks = []

    def timed(nsteps, fn):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps)]
        s.record()
        for i in range(nsteps):
            fn()
            marks[i].record()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        step_marks.append([round((s if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 2) for i in range(nsteps)])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- warm-up (eager), launch count per step, per-kernel CUDA-event profile of one eager pass
    W = max(args.warmup, 3)
    for _ in range(W):
        model.draw_noise(Bl, dev, out=noise_bufs)
        augment_views(devin)
        step(devin)
    l0 = _lib.launch_count()
    ops.PROFILE.clear()
    ops.PROFILE_ON = True
    prof_steps = 2
    for _ in range(prof_steps):
        model.draw_noise(Bl, dev, out=noise_bufs)
        augment_views(devin)
        step(devin)
    torch.cuda.synchronize()
    ops.PROFILE_ON = False
    launches = (_lib.launch_count() - l0) // prof_steps
    prof = ops.profile_summary()

    # ---- the step as ONE CUDA graph (forward, backward, NCCL collectives, fused optimizers); the vMF noise is
    #      redrawn in place before every replay by four launches outside the graph (host-side RNG counter)
    launch_mode = "eager"
    gstep = None
    phase_events = None
    if args.graph != "off":
        try:
            # the phase regions of ops._Prof become event-record NODES of the graph (external events): every replay
            # re-times them, so the per-phase breakdown below is measured inside the timed region itself
            ops.PROFILE.clear()
            ops.PROFILE_ON, ops.PROFILE_EXTERNAL = True, True
            step_s = torch.cuda.Event(enable_timing=True, external=True)
            step_e = torch.cuda.Event(enable_timing=True, external=True)

            def graphed():
                if ops.PROFILE_ON:              # external events can only be recorded under capture
                    step_s.record()
                out = step(devin, capturable=True)
                if ops.PROFILE_ON:
                    step_e.record()
                return out
            ops.PROFILE_ON = False
            gstep = GraphedStep(graphed, warmup=2, stream=torch.cuda.current_stream(),
                                on_capture=lambda on: setattr(ops, "PROFILE_ON", on))
            ops.PROFILE_ON, ops.PROFILE_EXTERNAL = False, False
            phase_events = dict(ops.PROFILE)
            launch_mode = "cuda_graph"
        except Exception as ex:  # noqa: BLE001
            if args.graph == "on":
                raise
            launch_mode = f"eager (graph capture failed: {type(ex).__name__})"
            gstep = None
            phase_events = None
            ops.PROFILE_ON, ops.PROFILE_EXTERNAL = False, False
            torch.cuda.synchronize()

    def run_value():
        model.draw_noise(Bl, dev, out=noise_bufs)
        augment_views(devin)
        if gstep is not None:
            gstep()
        else:
            step(devin)
    for _ in range(2):
        run_value()
    # ---- main number: device-resident inputs
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(args.steps, run_value)
    sampler.join(timeout=3)
    value = Bg * args.steps / (ms / 1e3)
    # per-phase device time of the LAST timed replay (event nodes inside the graph), max over ranks
    phases = None
    if phase_events:
        names = sorted(phase_events)
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in phase_events[k]) for k in names] + [step_s.elapsed_time(step_e)],
                         device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = {k: round(float(v), 3) for k, v in zip(names, t[:-1])}
        phases["graph_total"] = round(float(t[-1]), 3)
        phases["unattributed"] = round(float(t[-1]) - sum(phases[k] for k in names), 3)

    # ---- end-to-end: pinned host -> device copies every step + loss read back
    copy_stream = torch.cuda.Stream()
    bufs = [with_views({k: torch.empty_like(v, device=dev) for k, v in host.items()}) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    gsteps = [None, None]
    # The H2D prefetch of the next batch is started from INSIDE the step: an external event recorded by a node of the
    # graph at the start of the InfoNCE backward (long kernels; the launch-bound front of the step -- noise draws,
    # casts, small GEMMs, the embedding all-gathers -- is over by then).  Started at the step boundary instead, the
    # 67 MB / rank copy cost the 8-GPU step 1.6 ms (13.3 -> 14.9 ms) although it takes only 2.9 ms by itself.
    mid_ev = [torch.cuda.Event(external=True), torch.cuda.Event(external=True)] if not args.h2d_at_step_start else [None, None]
    if gstep is not None:
        try:
            for i in range(2):
                for k, v in host.items():
                    bufs[i][k].copy_(v)
                augment_views(bufs[i])
                fired = []

                def marker(name, i=i, fired=fired):
                    if name == "infonce_bwd" and not fired and mid_ev[i] is not None and torch.cuda.is_current_stream_capturing():
                        mid_ev[i].record()
                        fired.append(1)
                ops.ON_PHASE = marker
                try:
                    gsteps[i] = GraphedStep(lambda i=i: step(bufs[i], capturable=True), warmup=1, pool=gstep.pool(),
                                            stream=torch.cuda.current_stream())
                finally:
                    ops.ON_PHASE = None
                if not fired:
                    mid_ev[i] = None
        except Exception:  # noqa: BLE001
            gsteps = [None, None]
            ops.ON_PHASE = None
            torch.cuda.synchronize()

    copy_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    step_t0 = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    h2d_src = {"host": host}

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_ev[0].record(copy_stream)
            for k, v in h2d_src["host"].items():
                bufs[i][k].copy_(v, non_blocking=True)
            copy_ev[1].record(copy_stream)
            ready[i].record(copy_stream)
    state = {"i": 0, "losses": []}
    d2h = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h_ev = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_pending = [False, False]

    def e2e_step(copy=True, sync=True):
        # Every step: H2D of the NEXT batch on the copy stream (overlaps this step's compute), this step's
        # graph launch, an asynchronous D2H copy of its loss into pinned memory, and the host READ of the
        # previous step's loss (after its copy event) -- the usual one-step-lagged logging of a training loop,
        # so the host never idles the GPU while it waits for a scalar.
        i = state["i"]
        late = copy and gsteps[i] is not None and mid_ev[i] is not None
        if copy:
            torch.cuda.current_stream().wait_event(ready[i])
            if not late:
                copy_stream.wait_stream(torch.cuda.current_stream())   # next copy must not overwrite a buffer in use
                prefetch(i ^ 1)
        model.draw_noise(Bl, dev, out=noise_bufs)
        augment_views(bufs[i])
        step_t0[i].record()
        out = gsteps[i]() if gsteps[i] is not None else step(bufs[i])
        if late:
            # bufs[i ^ 1] was last read by step i-1, which precedes this step's marker in stream order
            copy_stream.wait_event(mid_ev[i])
            prefetch(i ^ 1)
        d2h[i].copy_(out.reshape(1), non_blocking=True)         # device->host read of the step's loss
        d2h_ev[i].record()
        d2h_pending[i] = True
        j = i ^ 1
        if sync and d2h_pending[j]:
            d2h_ev[j].synchronize()
            state["losses"].append(float(d2h[j][0]))
            d2h_pending[j] = False
        state["i"] = i ^ 1

    def e2e_drain():
        for j in range(2):
            if d2h_pending[j]:
                d2h_ev[j].synchronize()
                state["losses"].append(float(d2h[j][0]))
                d2h_pending[j] = False
    prefetch(0)
    e2e_step()
    ms_e2e = timed(args.steps, e2e_step)
    e2e_drain()
    e2e_value = Bg * args.steps / (ms_e2e / 1e3)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    h2d_copy_ms = max_over_ranks(copy_ev[0].elapsed_time(copy_ev[1]))     # the last in-loop copy, device-timed
    # how far into its step the last prefetch started (the graph's marker event releases it)
    h2d_start_ms = step_t0[state["i"] ^ 1].elapsed_time(copy_ev[0])
    e2e_diag = None
    if args.e2e_diag:
        # where does the end-to-end arm lose time against the device-resident one?  Same loop without the H2D copies,
        # and with the copies but without the per-step host read of the loss.
        ms_nocopy = timed(args.steps, lambda: e2e_step(copy=False))
        e2e_drain()
        prefetch(state["i"])
        ms_nosync = timed(args.steps, lambda: e2e_step(sync=False))
        e2e_drain()
        e2e_diag = {"ms_per_step_no_h2d": ms_nocopy / args.steps, "ms_per_step_no_host_read": ms_nosync / args.steps,
                    "h2d_copy_ms_in_loop": h2d_copy_ms, "h2d_start_ms_into_step": h2d_start_ms,
                    "numa_bound_cpus": len(local_cpus) if local_cpus else None}
        if local_cpus:      # the same loop fed from pinned buffers allocated BEFORE the NUMA binding
            h2d_src["host"] = host_unbound
            prefetch(state["i"])
            ms_unbound = timed(args.steps, e2e_step)
            e2e_drain()
            e2e_diag["ms_per_step_unbound_pinned"] = ms_unbound / args.steps
            e2e_diag["h2d_copy_ms_unbound"] = max_over_ranks(copy_ev[0].elapsed_time(copy_ev[1]))
            h2d_src["host"] = host

    if local_cpus:
        os.sched_setaffinity(0, all_cpus)       # the CPU baseline arm below uses every host thread
    if rank != 0:
        return
    pk = peaks()
    # rooflines of the two K2 kernels that carry ~80 % of the step, measured live (CUDA events on the launching stream
    # around each step's launches of the phase, which are dealt over two streams: fork to join):
    #   forward  rowcol_sum_tc4_kernel: 8 launches per step = 4 cross blocks (2 Bl Bg D FLOP each, row + column sums,
    #            the exp(s - shift) blocks written to HBM for the backward) + 4 symmetric half-window blocks of the two
    #            critic calls whose no-grad diagnostics the reference logs (Bl Bg D each)
    #   backward infonce_bwd_e_kernel: 8 launches per step (4 critic calls x 2 gradient directions), 2 Bl Bg D FLOP each,
    #            ONE product from the stored probabilities (executed = algorithmic FLOPs)
    # `roofline` is the one with the larger share of the step.  DRAM traffic and tensor-pipe activity come from the
    # committed ncu capture of the same kernel at this shape (profiles/r02_traffic.json), never from literals here.
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    caps = {}
    if os.path.isfile(tpath):
        try:
            caps = json.load(open(tpath))
        except Exception:  # noqa: BLE001
            caps = {}
    same_shape = world == 1 and Bg == 65536 and prec == "bf16"

    def k2_roof(phase, kname, launches_per_step, flops_per_step, what):
        if phase not in prof:
            return None
        _, tot_ms = prof[phase]                  # total over the profiled steps (a phase may be several timed regions)
        n = prof_steps * launches_per_step
        ach = flops_per_step * prof_steps / (tot_ms * 1e-3) / 1e12
        cap = caps.get(kname, {})
        return {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_sust"], "traffic": cap.get("dram_bytes") if same_shape else None, "launches": n,
                "avg_ms": tot_ms / n, "peak_source": pk["src"] + " sustained bf16 (kernel timed inside a long step)",
                "share_of_step": (tot_ms / prof_steps) / (ms / args.steps),
                "algorithmic_flops_per_step": flops_per_step, "what": what,
                "timed_in": f"{prof_steps} eager steps after warm-up (CUDA events on the launching stream around each step's "
                            f"{launches_per_step} launches, fork to join of the two streams they are dealt over)",
                "traffic_source": ("bytes/launch, ncu --set full capture of this kernel at this shape "
                                   "(profiles/r02_traffic.json)") if same_shape and cap else None,
                "executed_over_algorithmic_flops": cap.get("executed_over_algorithmic_flops"),
                "tensor_pipe_active_pct_ncu": cap.get("tensor_pipe_active_pct") if same_shape else None}
    stored_e = bool(ops._STORE_E) and prec == "bf16" and EMB in (256, 512)
    bwd_kernel = "infonce_bwd_tc6_kernel"
    if stored_e:
        bwd_kernel = "infonce_bwd_e_kernel"
        if world > 1 and not ops._DP_RS:
            bwd_kernel = "infonce_bwd_e_kernel (local rows) + infonce_bwd_tc6_kernel (gathered columns)"
    roof_fwd = k2_roof("rowlse_x4", "rowcol_sum_tc4_kernel", 8, 12.0 * Bl * Bg * EMB,
                       "4 cross blocks (row + column sums" + (", exp(s - shift) kept in HBM as bf16" if stored_e else "") +
                       ") + 4 symmetric half-window blocks per step")
    roof_bwd = k2_roof("infonce_bwd", bwd_kernel, 8, 16.0 * Bl * Bg * EMB,
                       "4 critic calls x 2 gradient directions per step" +
                       (", one product each from the stored probabilities" if stored_e else ", S recomputed per launch") +
                       ("; the view-1 partial gradients of the ranks are summed by one NCCL reduce-scatter per call"
                        if stored_e and world > 1 and ops._DP_RS else ""))
    roof = None
    cands = [r for r in (roof_fwd, roof_bwd) if r]
    if cands:
        roof = dict(max(cands, key=lambda r: r["share_of_step"]))
        roof["other_kernels_ms_per_step"] = {k: v[1] / prof_steps for k, v in prof.items()}
    # rooflines of the other two kernel families, measured live (CUDA events, kernels timed alone -> burst peak):
    #   K1 grouped MLP GEMM at the C5 layer shapes (2 groups, M = 2 x per-GPU batch), K3 evidence fusion + EDL loss in
    #   the mode training uses (gradient + conflict term) at the C4 shape B = 2^22, V = 4, C = 42
    # algorithmic FLOPs of one step: K2 = 2 critic calls with the no-grad diagnostics (8 B^2 D each) + the 2
    # specific-critic calls whose diagnostics the reference discards (6 B^2 D each: cross block fwd 2 + bwd 4);
    # K1 = 48.2 MFLOP/sample; ortho Grams forward only (lmd = 0: logged, weight exactly zero) 8 D^2 per sample
    step_flops = 28.0 * Bg * Bg * EMB + Bg * (48.2e6 + 8.0 * EMB * EMB) if Bg else 0
    cpu = None
    if not args.no_cpu_baseline:
        cores = host_threads()
        torch.set_float32_matmul_precision("highest")
        cval, cdt = time_cpu(args.cpu_batch, 2, 1)
        cpu = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 full steps of the oracle port at batch {args.cpu_batch} (reference cost grows ~B^2; B=65536 needs 64 GiB per logits temp)"}
        try:        # the reference imports with torch.set_float32_matmul_precision('medium') (models/dmvae.py:9)
            torch.set_float32_matmul_precision("medium")
            mval, _ = time_cpu(args.cpu_batch, 1, 1)
            cpu["value_medium_precision"] = mval
        except Exception:  # noqa: BLE001
            pass
        finally:
            torch.set_float32_matmul_precision("highest")
        # context only: the same oracle port run with torch-eager ATen kernels ON THIS GPU at the largest batch
        # whose [2B,2B] logits it can hold comfortably
        try:
            cpu["gpu_eager_context"] = gpu_eager_context(dev, 8192)
        except Exception as ex:  # noqa: BLE001
            cpu["gpu_eager_context"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:160]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if prec == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "C5 DisentangledSSL 2x1024-d, hidden 512, embed 512, T=0.07 + 3-head evidential probe (C=10)",
                   "global_batch": Bg, "per_gpu_batch": Bl, "parallelism": f"dp{world}",
                   "infonce_backward": ("one product per gradient from exp(s - shift) kept in HBM by the forward "
                                        f"({4 * 2 * ((Bl + 255) // 256) * 4 * ((Bg + 255) // 256) * 16384 / 2 ** 30:.1f} GiB/step/GPU)"
                                        if stored_e else "S recomputed"),
                   "l2": "inputs (1 GiB/step/GPU at dp1) and embeddings are larger than the 126 MB L2",
                   "noise": "vMF noise and the augmented views v1, v2 drawn on device every step (dmf_vmf_draw, dmf_augment)", "launch": launch_mode},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
        "step_tflops_algorithmic": step_flops / 1e12,
        "step_frac_of_sustained_bf16": step_flops / (ms / args.steps * 1e-3) / 1e12 / (pk["tf_sust"] * world),
        "roofline": roof, "roofline_k2_fwd": roof_fwd, "roofline_k2_bwd": roof_bwd, "roofline_k1": roof_k1, "roofline_k3": roof_k3, "cpu_baseline": cpu,
        "loss_check": lcheck, "phase_ms": phases, "e2e_diag": e2e_diag,
        "step_ms": {"value": step_marks[0], "e2e": step_marks[1]},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="GLOBAL batch (sharded over ranks)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c5", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json config: c5 (default) = the headline workload; c1-c4 = the small launch-bound configs "
                         "(single GPU; samples/s + launches/step beside the CPU port at the same batch)")
    ap.add_argument("--e2e-diag", action="store_true", help="extra end-to-end loops without H2D / without the host read")
    ap.add_argument("--nccl-priority", default="high", choices=["high", "normal"],
                    help="priority of the NCCL stream (high: collective CTAs take the next free SM under the InfoNCE tiles)")
    ap.add_argument("--h2d-at-step-start", action="store_true",
                    help="start the H2D prefetch at the step boundary instead of at the graph's mid-step marker (A/B)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="N > 1: do not bind each rank to the CPUs of its GPU's NUMA node while it allocates pinned memory")
    ap.add_argument("--no-kernel-rooflines", action="store_true", help="skip the live K1 / K3 roofline measurements")
    ap.add_argument("--no-loss-check", dest="loss_check", action="store_false",
                    help="skip the cross-N loss / gradient-norm check step")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step as one CUDA graph (auto: fall back to eager launches if capture fails)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config != "c5":
        if rank == 0:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream(device=0)):      # see dp.GraphedStep: capture needs a non-default stream
                run_small_config(args)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        from disentagled_multimodal_fusion_b200.dp import init_process_group_nccl
        init_process_group_nccl(torch.device("cuda", local_rank), high_priority=args.nccl_priority == "high")
    try:
        # the whole run lives on one non-default stream (see dp.GraphedStep: CUDA-graph capture of the backward
        # pass needs every autograd leaf to have been touched on a capturable stream only)
        main_stream = torch.cuda.Stream(device=local_rank)
        with torch.cuda.stream(main_stream):
            run_gpu(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
