"""torch.autograd.Function wrappers over the C ABI (include/dmf_b200.h).

Every function here launches hand-written sm_100a kernels from libdmf_b200.so on the current
torch CUDA stream; torch supplies device memory, streams and (for data parallelism) the NCCL
process group -- nothing else.  ``precision`` is 'fp32' (FFMA path, 1e-5 parity) or 'bf16'
(tcgen05 tensor-core path, 2e-2 parity).
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from ._lib import lib, check, ptr, stream

Tensor = torch.Tensor


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ---- optional per-kernel timing with CUDA events on the launching stream (used by bench.py)
PROFILE_ON = False
PROFILE_EXTERNAL = False     # events recorded as nodes of a CUDA graph under capture (timed on every replay)
PROFILE: dict = {}
ON_PHASE = None              # optional callback(name) at the start of every phase region (e.g. record a marker event
                             # into the CUDA graph under capture: bench.py starts its H2D prefetch at such a marker)


class _Prof:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if ON_PHASE is not None:
            ON_PHASE(self.name)
        if PROFILE_ON:
            self.s = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
            self.e = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if PROFILE_ON:
            self.e.record()
            PROFILE.setdefault(self.name, []).append((self.s, self.e))
        return False


def profile_summary():
    """{name: (number of timed regions, total ms)}; call after a device synchronize."""
    return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in PROFILE.items()}


# ----------------------------------------------------------------------------------------
# low-level helpers (no autograd)
# ----------------------------------------------------------------------------------------
def gemm_f32(descs: Sequence[dict], epilogue: int) -> None:
    """descs: dicts with A,a_rs,a_cs,B,b_rs,b_cs,C,ldc,M,N,K and optional bias,aux,ldaux,rowsum_a,accumulate."""
    L.require_device()
    for base in range(0, len(descs), 64):
        chunk = descs[base:base + 64]
        arr = (L.GemmDesc * len(chunk))()
        for i, d in enumerate(chunk):
            g = arr[i]
            g.A, g.a_rs, g.a_cs = ptr(d["A"]), d["a_rs"], d["a_cs"]
            g.B, g.b_rs, g.b_cs = ptr(d["B"]), d["b_rs"], d["b_cs"]
            g.C, g.ldc = ptr(d["C"]), d["ldc"]
            g.bias = ptr(d.get("bias"))
            g.aux, g.ldaux = ptr(d.get("aux")), d.get("ldaux", 0)
            g.rowsum_a = ptr(d.get("rowsum_a"))
            g.M, g.N, g.K = d["M"], d["N"], d["K"]
            g.accumulate = d.get("accumulate", 0)
        check(lib.dmf_grouped_gemm_f32(arr, len(chunk), epilogue, stream()))


def gemm_tc(descs: Sequence[dict], epilogue: int) -> None:
    """descs: dicts with A,lda,B,ldb,M,N,K and out_f32/ldo_f32, out_bf16/ldo_bf16, bias, mask/ldmask."""
    L.require_device()
    arr = (L.TcGemmDesc * len(descs))()
    for i, d in enumerate(descs):
        g = arr[i]
        g.A, g.lda, g.B, g.ldb = ptr(d["A"]), d["lda"], ptr(d["B"]), d["ldb"]
        g.out_f32, g.ldo_f32 = ptr(d.get("out_f32")), d.get("ldo_f32", 0)
        g.out_bf16, g.ldo_bf16 = ptr(d.get("out_bf16")), d.get("ldo_bf16", 0)
        g.bias = ptr(d.get("bias"))
        g.mask_bf16, g.ldmask = ptr(d.get("mask")), d.get("ldmask", 0)
        g.M, g.N, g.K = d["M"], d["N"], d["K"]
        g.out_bf16_t, g.ldo_t = ptr(d.get("out_t")), d.get("ldo_t", 0)
        g.split_k = d.get("split_k", 1)
        g.mn_major = d.get("mn_major", 0)
    check(lib.dmf_grouped_gemm_bf16_tc(arr, len(descs), epilogue, stream()))


def head_gemm(descs: Sequence[dict], head: int) -> None:
    """Last encoder layer with the row head in its epilogue (dmf_head_gemm_bf16): head 0 = row-normalise, 1 = vMF sample.
    Each dict: A, W (bf16, row-major), bias, M, N, K and any of pre_f32 / pre_bf16 / out_f32 / out_bf16 / inv_norm /
    noise_w / noise_v / eps (tensors; leading dimensions are taken from their strides)."""
    arr = (L.HeadGemmDesc * len(descs))()
    for i, d in enumerate(descs):
        g = arr[i]
        g.A, g.lda, g.W, g.ldw, g.bias = ptr(d["A"]), d["A"].stride(0), ptr(d["W"]), d["W"].stride(0), ptr(d["bias"])
        for k in ("pre_f32", "pre_bf16", "out_f32", "out_bf16"):
            t = d.get(k)
            setattr(g, k, ptr(t))
            setattr(g, "ld_" + k, t.stride(0) if t is not None else 0)
        g.inv_norm, g.noise_w, g.noise_v = ptr(d.get("inv_norm")), ptr(d.get("noise_w")), ptr(d.get("noise_v"))
        g.eps = float(d.get("eps", 1e-12))
        g.M, g.N, g.K = int(d["M"]), int(d["N"]), int(d["K"])
    check(lib.dmf_head_gemm_bf16(arr, len(descs), int(head), stream()))


def head_fusable(last_shapes, out_bt=None) -> bool:
    """Can the row head (row-normalise / vMF sample) run in the epilogue of the last layer?  ``last_shapes`` = [out, in]
    of every group's last Linear.  One CTA pair must own whole rows: widths 256 / 512 only."""
    # Opt-in (DMF_FUSE_HEADS=1).  Measured at C5 (profiles/r02_bench_v9_fused_heads.log): parity green, but the step is
    # 64.7 ms against 60.2 ms with the separate head kernels -- a whole 512-wide row fills the 512 TMEM columns of its CTA,
    # so the two-pass epilogue (1 MB of stores per 128-row tile) cannot overlap the next tile's main loop the way the
    # double-buffered 256-column pair GEMM does, and what the fusion saves is one re-read of X (0.04 ms per group).
    if os.environ.get("DMF_FUSE_HEADS", "0") != "1":
        return False
    if out_bt is not None and any(t is not None for t in out_bt):
        return False
    return all(int(n) in (256, 512) for n, k in last_shapes)


def wgrad_mn_ok(n_out: int, k_in: int) -> bool:
    """True when the wgrad of a Linear(k_in -> n_out) runs on the CTA-pair kernel with MN-major operands, i.e. reads the
    row-major bf16 activations dY [batch, n_out] and X [batch, k_in] as they are: no transposed copies are written
    anywhere.  (``DMF_WGRAD_T=1`` forces the older transposed-copy path, for A/B timing.)"""
    return n_out >= 512 and k_in >= 128 and n_out % 8 == 0 and not _WGRAD_T


_WGRAD_T = bool(os.environ.get("DMF_WGRAD_T"))


def cast_bf16(src: Tensor, dst: Optional[Tensor] = None, ldd: Optional[int] = None) -> Tensor:
    """fp32 [R,C] (row stride src.stride(0)) -> bf16; dst may be a column slice of a wider buffer."""
    R, Cc = src.shape
    if dst is None:
        dst = torch.empty(R, Cc, dtype=torch.bfloat16, device=src.device)
    check(lib.dmf_cast_f32_to_bf16(ptr(src), src.stride(0), ptr(dst), ldd or dst.stride(0), R, Cc, stream()))
    return dst


def cast_dual_bf16(src: Tensor, dst: Optional[Tensor], ldd: int, dstT: Optional[Tensor], ldt: int,
                   colsum_out: Optional[Tensor] = None) -> None:
    """One pass over fp32 ``src`` [R,C]: bf16 copy into ``dst`` (row pitch ldd), transposed bf16 copy into
    ``dstT`` (row pitch ldt) and optionally column sums accumulated into ``colsum_out`` (caller zeroes)."""
    L.require_device()
    src = _f32c(src)
    R, Cc = src.shape
    check(lib.dmf_cast_dual_bf16(ptr(src), src.stride(0), ptr(dst), ldd, ptr(dstT), ldt, ptr(colsum_out), R, Cc, stream()))


def cast_transpose_bf16(src: Tensor) -> Tensor:
    R, Cc = src.shape
    Rp = (R + 7) // 8 * 8                       # TMA needs 16-byte row pitch
    dst = torch.empty(Cc, Rp, dtype=torch.bfloat16, device=src.device)
    if Rp != R:
        dst[:, R:].zero_()
    check(lib.dmf_cast_transpose_f32_to_bf16(ptr(src), src.stride(0), ptr(dst), Rp, R, Cc, stream()))
    return dst


def transpose_bf16(src: Tensor) -> Tensor:
    R, Cc = src.shape
    Rp = (R + 7) // 8 * 8
    dst = torch.empty(Cc, Rp, dtype=torch.bfloat16, device=src.device)
    if Rp != R:
        dst[:, R:].zero_()
    check(lib.dmf_transpose_bf16(ptr(src), src.stride(0), ptr(dst), Rp, R, Cc, stream()))
    return dst


def colsum(x: Tensor, out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    M, N = x.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x.device)
    check(lib.dmf_colsum_f32(ptr(x), x.stride(0), M, N, ptr(out), 1 if accumulate else 0, stream()))
    return out


def _grad_slot(p: Tensor) -> Optional[Tensor]:
    """The pre-allocated fp32 .grad of a leaf parameter (dp.FlatParams points every .grad into one flat buffer),
    or None.  Kernels that can accumulate (red.add) write weight / bias gradients straight into it and the
    autograd Function returns None for that input: no temporary, no zero-fill, no AccumulateGrad add."""
    g = getattr(p, "grad", None)
    if g is None or not p.is_leaf or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


def rng_counter(device) -> Tensor:
    """A device-resident 64-bit draw counter for the ``ctr=`` argument of ``vmf_draw`` / ``augment``: the draws are keyed
    by (seed, counter), and ``counter_add`` bumps it ON THE DEVICE -- so a training step captured once into a CUDA
    graph draws fresh noise on every replay."""
    return torch.zeros(1, dtype=torch.int64, device=device)


def counter_add(ctr: Tensor, inc: int = 1) -> None:
    L.require_device()
    check(lib.dmf_counter_add(ptr(ctr), int(inc), stream()))


def vmf_draw(rows: int, D: int, kappa: float, seed: int, offset: int, device, out=None, ctr: Optional[Tensor] = None
             ) -> Tuple[Tensor, Tensor]:
    """Device-side vMF noise (distribution-equal to the reference sampler, not stream-equal).  ``out`` = (w, v)
    pre-allocated buffers to draw into (fixed addresses for CUDA-graph replay); ``ctr`` = device draw counter."""
    L.require_device()
    if out is None:
        w = torch.empty(rows, 1, dtype=torch.float32, device=device)
        v = torch.empty(rows, D - 1, dtype=torch.float32, device=device)
    else:
        w, v = out
    if ctr is not None:
        check(lib.dmf_vmf_draw_ctr(ptr(w), ptr(v), rows, D, float(kappa), seed, offset, ptr(ctr), stream()))
    else:
        check(lib.dmf_vmf_draw(ptr(w), ptr(v), rows, D, float(kappa), seed, offset, stream()))
    return w, v


def augment(x: Tensor, seed: int, offset: int = 0, noise_scale: float = 0.01, drop_scale: int = 10,
            return_choice: bool = False, out: Optional[Tensor] = None, ctr: Optional[Tensor] = None):
    """Device-side ``augment_data`` (utils.py:118-151): per row noise / random column drop / identity with
    probability 1/3 each.  Distribution-equal to the reference's host loop, not stream-equal."""
    L.require_device()
    x = _f32c(x)
    B, D = x.shape
    y = torch.empty_like(x) if out is None else out        # ``out``: fixed buffer refilled in place (CUDA-graph replays)
    ch = torch.empty(B, dtype=torch.int32, device=x.device) if return_choice else None
    if ctr is not None:
        check(lib.dmf_augment_ctr(ptr(x), x.stride(0), ptr(y), y.stride(0), B, D, float(noise_scale), int(drop_scale),
                                  int(seed), int(offset), ptr(ch), ptr(ctr), stream()))
    else:
        check(lib.dmf_augment(ptr(x), x.stride(0), ptr(y), y.stride(0), B, D, float(noise_scale), int(drop_scale), int(seed),
                              int(offset), ptr(ch), stream()))
    return (y, ch) if return_choice else y


# ----------------------------------------------------------------------------------------
# K1  grouped MLP   (models/classifiers.py:16-48, 469-502)
# ----------------------------------------------------------------------------------------
class _GroupedMLP(torch.autograd.Function):
    """forward(cfg, *tensors): tensors = G inputs, G "extra" inputs (None or a tensor whose columns
    are appended to the input: layer-0 input = [x | extra], gradient flows to ``extra`` only), then
    per group L weights, then per group L biases.
    cfg = (G, L, final, precision, dropout_masks) with final in {'none', 'evidence'}.
    One kernel launch per layer covers all groups."""

    @staticmethod
    def forward(ctx, cfg, *tensors):
        L.require_device()
        G, NL, final, precision, masks, opts = cfg
        xs = [t for t in tensors[:G]]
        extras = [t for t in tensors[G:2 * G]]
        o = 2 * G
        Ws = [list(tensors[o + g * NL: o + (g + 1) * NL]) for g in range(G)]
        bs = [list(tensors[o + G * NL + g * NL: o + G * NL + (g + 1) * NL]) for g in range(G)]
        ctx.cfg = cfg
        ctx.in_needs_grad = [bool(x.requires_grad) and extras[g] is None for g, x in enumerate(xs)]
        ctx.extra_cols = [0 if e is None else e.shape[1] for e in extras]
        ctx.extra_needs_grad = [e is not None and bool(e.requires_grad) for e in extras]
        if precision == "bf16" and masks is not None and any(m is not None and any(x is not None for x in m) for m in masks):
            # the tensor-core epilogues carry no dropout mask: refuse loudly instead of training without dropout
            raise NotImplementedError("bf16 grouped MLP with dropout masks: use precision='fp32' or dropout=0 "
                                      "(the reference configs run the DMVAE / DSSL encoders with dropout 0)")
        with _Prof("mlp_fwd"):
            if precision == "bf16":
                outs, saved = _GroupedMLP._fwd_bf16(xs, extras, Ws, bs, G, NL, final, opts or {})
            else:
                xs = [x if e is None else torch.cat([_f32c(x), _f32c(e)], dim=1) for x, e in zip(xs, extras)]
                outs, saved = _GroupedMLP._fwd_f32(xs, Ws, bs, G, NL, final, masks)
        ctx.saved = saved
        ctx.Ws, ctx.bs = Ws, bs
        if precision == "bf16" and saved.get("head") is not None:
            ctx.mark_non_differentiable(*outs[2 * G:])          # the bf16 copies of the head values
        return tuple(outs)

    # ---------------- fp32 (FFMA) path
    @staticmethod
    def _fwd_f32(xs, Ws, bs, G, NL, final, masks):
        xs = [_f32c(x) for x in xs]
        acts = [[x] for x in xs]
        pre = [None] * G
        for l in range(NL):
            last = l == NL - 1
            descs = []
            for g in range(G):
                A = acts[g][-1]
                W = Ws[g][l]
                M, K = A.shape
                N = W.shape[0]
                out = torch.empty(M, N, dtype=torch.float32, device=A.device)
                d = dict(A=A, a_rs=A.stride(0), a_cs=1, B=W, b_rs=1, b_cs=W.stride(0), C=out, ldc=N,
                         bias=bs[g][l], M=M, N=N, K=K)
                if last and final == "evidence":
                    pre[g] = torch.empty(M, N, dtype=torch.float32, device=A.device)
                    d.update(aux=pre[g], ldaux=N)
                descs.append(d)
                acts[g].append(out)
            epi = (L.EPI_BIAS_EVIDENCE if final == "evidence" else L.EPI_BIAS) if last else L.EPI_BIAS_RELU
            gemm_f32(descs, epi)
            if not last and masks is not None:
                for g in range(G):
                    if masks[g] is not None and masks[g][l] is not None:
                        acts[g][-1].mul_(masks[g][l])       # inverted-dropout mask (train mode only)
        outs = [acts[g][-1] for g in range(G)]
        return outs, dict(acts=acts, pre=pre, masks=masks)

    @staticmethod
    def _bwd_f32(ctx, grads):
        G, NL, final, precision, masks, opts = ctx.cfg
        acts, pre = ctx.saved["acts"], ctx.saved["pre"]
        Ws = ctx.Ws
        dev = acts[0][0].device
        dYs = []
        for g in range(G):
            dy = grads[g]
            if dy is None:
                dy = torch.zeros_like(acts[g][-1])
            dy = _f32c(dy)
            if final == "evidence":
                dh = torch.empty_like(dy)
                check(lib.dmf_evidence_bwd(ptr(pre[g]), ptr(acts[g][-1]), ptr(dy), ptr(dh), dy.numel(), stream()))
                dy = dh
            dYs.append(dy)
        dWs = [[None] * NL for _ in range(G)]
        dbs = [[None] * NL for _ in range(G)]
        dxs = [None] * G
        for l in range(NL - 1, -1, -1):
            wdescs, ddescs, partials = [], [], []
            new_dY = [None] * G
            for g in range(G):
                X = acts[g][l]
                dY = dYs[g]
                M, K = X.shape
                N = dY.shape[1]
                # dW[n,k] = sum_m dY[m,n] X[m,k];  db[n] = sum_m dY[m,n]  (row sums of the A operand).
                # The output is tiny and the contraction runs over the batch: split the batch into chunks
                # (one descriptor each, partial results reduced by a column sum) so the grid fills the GPU.
                chunks = max(1, min(32, M // 2048))
                if chunks == 1:
                    dW = torch.empty(N, K, dtype=torch.float32, device=dev)
                    db = torch.empty(N, dtype=torch.float32, device=dev)
                    wdescs.append(dict(A=dY, a_rs=1, a_cs=dY.stride(0), B=X, b_rs=X.stride(0), b_cs=1, C=dW, ldc=K,
                                       rowsum_a=db, M=N, N=K, K=M))
                    dWs[g][l], dbs[g][l] = dW, db
                else:
                    rows = (M + chunks - 1) // chunks
                    chunks = (M + rows - 1) // rows
                    dWp = torch.empty(chunks, N * K, dtype=torch.float32, device=dev)
                    dbp = torch.empty(chunks, N, dtype=torch.float32, device=dev)
                    for c in range(chunks):
                        r0 = c * rows
                        rc = min(rows, M - r0)
                        wdescs.append(dict(A=dY[r0:], a_rs=1, a_cs=dY.stride(0), B=X[r0:], b_rs=X.stride(0), b_cs=1,
                                           C=dWp[c], ldc=K, rowsum_a=dbp[c], M=N, N=K, K=rc))
                    partials.append((g, l, dWp, dbp, N, K))
                if l > 0 or ctx.in_needs_grad[g]:
                    W = Ws[g][l]
                    dX = torch.empty(M, K, dtype=torch.float32, device=dev)
                    d = dict(A=dY, a_rs=dY.stride(0), a_cs=1, B=W, b_rs=W.stride(0), b_cs=1, C=dX, ldc=K, M=M, N=K, K=N)
                    if l > 0:
                        d.update(aux=X, ldaux=X.stride(0))     # ReLU backward: mask by the saved activation
                    ddescs.append((g, d))
                    new_dY[g] = dX
                elif l == 0 and ctx.extra_needs_grad[g]:
                    W = Ws[g][l]
                    De = ctx.extra_cols[g]
                    Wv = W[:, K - De:]                         # only the appended columns need a gradient
                    dX = torch.empty(M, De, dtype=torch.float32, device=dev)
                    ddescs.append((g, dict(A=dY, a_rs=dY.stride(0), a_cs=1, B=Wv, b_rs=W.stride(0), b_cs=1, C=dX,
                                           ldc=De, M=M, N=De, K=N)))
                    new_dY[g] = dX
            gemm_f32(wdescs, L.EPI_NONE)
            for (g, ll, dWp, dbp, N, K) in partials:
                ws, bsl = _grad_slot(Ws[g][ll]), _grad_slot(ctx.bs[g][ll])
                if ws is not None:
                    colsum(dWp, ws.view(-1), accumulate=True)
                else:
                    dWs[g][ll] = colsum(dWp).view(N, K)
                if bsl is not None:
                    colsum(dbp, bsl, accumulate=True)
                else:
                    dbs[g][ll] = colsum(dbp)
            if ddescs:
                if l > 0:
                    gemm_f32([d for _, d in ddescs], L.EPI_RELU_MASK)
                    if masks is not None:
                        for g, _ in ddescs:
                            if masks[g] is not None and masks[g][l - 1] is not None:
                                new_dY[g].mul_(masks[g][l - 1])
                else:
                    gemm_f32([d for _, d in ddescs], L.EPI_NONE)
            if l == 0:
                dxs = new_dY
            else:
                dYs = new_dY
        return dxs, dWs, dbs

    # ---------------- bf16 tensor-core path
    # Every activation / gradient that a later wgrad needs as a K-major operand (K = batch) is written
    # TRANSPOSED by the epilogue of the GEMM that produces it (out_t), so no transpose pass runs over them.
    @staticmethod
    def _fwd_bf16(xs, extras, Ws, bs, G, NL, final, opts):
        if final == "evidence":
            raise L.DmfError("bf16 grouped MLP: evidence epilogue is only built for the fp32 path")
        dev = xs[0].device
        need_t = torch.is_grad_enabled() and any(w.requires_grad for ws in Ws for w in ws)
        # layers whose wgrad reads the row-major activations MN-major need no transposed copy of their input
        mn = [all(wgrad_mn_ok(*Ws[g][l].shape) for g in range(G)) for l in range(NL)]
        xTs = opts.get("xTs") or [None] * G
        out_b = opts.get("out_bf16") or [None] * G
        out_bt = opts.get("out_bf16T") or [None] * G
        a0, a0T = [], []
        for g, (x, ex) in enumerate(zip(xs, extras)):
            if ex is not None:
                # x is a bf16 buffer [M, d + De] whose first d columns are filled; the tail columns hold
                # bf16(extra) -- either written already by the producing GEMM's epilogue or cast here
                De = ex.shape[1]
                if x.dtype != torch.bfloat16 or x.shape[1] < De:
                    raise L.DmfError("bf16 grouped MLP: with `extra`, pass the pre-cast bf16 concat buffer as x")
                if not opts.get("extras_prefilled", False):
                    d0 = x.shape[1] - De
                    exc = _f32c(ex)
                    check(lib.dmf_cast_f32_to_bf16(ptr(exc), De, x.data_ptr() + 2 * d0, x.stride(0), x.shape[0], De, stream()))
                    xTs[g] = None
                a0.append(x)
                a0T.append(xTs[g])
            elif x.dtype == torch.bfloat16:
                a0.append(x)
                a0T.append(xTs[g])
            else:
                x = _f32c(x)
                M, K = x.shape
                Kp, Mp = (K + 7) // 8 * 8, (M + 7) // 8 * 8
                buf = torch.empty(M, Kp, dtype=torch.bfloat16, device=dev)
                bufT = torch.empty(K, Mp, dtype=torch.bfloat16, device=dev) if (need_t and not mn[0]) else None
                check(lib.dmf_cast_dual_bf16(ptr(x), x.stride(0), ptr(buf), Kp, ptr(bufT), Mp, 0, M, K, stream()))
                a0.append(buf[:, :K] if Kp != K else buf)
                a0T.append(bufT)
        acts = [[a] for a in a0]
        actTs = [[t] for t in a0T]
        Wb = [[None] * NL for _ in range(G)]
        outs = []
        head = opts.get("head")
        if head is not None and not head_fusable([Ws[g][-1].shape for g in range(G)], out_bt):
            raise L.DmfError("grouped MLP: the fused row head needs an output width of 256 or 512 and no "
                             "transposed output copy (check ops.head_fusable first)")
        head_saved, head_out, head_outb = [], [], []
        for l in range(NL):
            last = l == NL - 1
            descs = []
            for g in range(G):
                A = acts[g][-1]
                W = Ws[g][l]
                N, K = W.shape
                Kp = (K + 7) // 8 * 8
                wb = torch.empty(N, Kp, dtype=torch.bfloat16, device=dev)
                cast_bf16(W, wb, Kp)
                Wb[g][l] = wb
                M = A.shape[0]
                Mp = (M + 7) // 8 * 8
                d = dict(A=A, lda=A.stride(0), B=wb, ldb=Kp, bias=bs[g][l], M=M, N=N, K=K)
                if last and head is not None:
                    # the row head runs in this layer's epilogue (dmf_head_gemm_bf16): X itself (fp32: ortho term, head
                    # backward; bf16: conditioning columns of the private encoders), the head value fp32 + bf16
                    o = torch.empty(M, N, dtype=torch.float32, device=dev)
                    y = torch.empty(M, N, dtype=torch.float32, device=dev)
                    yb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
                    hd = dict(A=A, W=wb, bias=bs[g][l], M=M, N=N, K=K, pre_f32=o, pre_bf16=out_b[g], out_f32=y, out_bf16=yb)
                    if head["kind"] == 0:
                        inv = torch.empty(M, dtype=torch.float32, device=dev)
                        hd.update(inv_norm=inv, eps=head.get("eps", 1e-12))
                        head_saved.append((y, inv))
                    else:
                        w_, v_ = head["noise"][g]
                        hd.update(noise_w=_f32c(w_), noise_v=_f32c(v_))
                        head_saved.append((o, hd["noise_w"], hd["noise_v"]))
                    outs.append(o)
                    head_out.append(y)
                    head_outb.append(yb)
                    descs.append(hd)
                    continue
                if last:
                    o = torch.empty(M, N, dtype=torch.float32, device=dev)
                    d.update(out_f32=o, ldo_f32=N)
                    if out_b[g] is not None:
                        d.update(out_bf16=out_b[g], ldo_bf16=out_b[g].stride(0))
                    if out_bt[g] is not None:
                        d.update(out_t=out_bt[g], ldo_t=out_bt[g].stride(0))
                    outs.append(o)
                else:
                    Np = (N + 7) // 8 * 8
                    o = torch.empty(M, Np, dtype=torch.bfloat16, device=dev)
                    d.update(out_bf16=o, ldo_bf16=Np)
                    acts[g].append(o[:, :N] if Np != N else o)
                    if need_t and not mn[l + 1]:
                        oT = torch.empty(N, Mp, dtype=torch.bfloat16, device=dev)
                        d.update(out_t=oT, ldo_t=Mp)
                        actTs[g].append(oT)
                    else:
                        actTs[g].append(None)
                descs.append(d)
            if last and head is not None:
                head_gemm(descs, head["kind"])
            else:
                gemm_tc(descs, L.EPI_BIAS if (last and final != "relu") else L.EPI_BIAS_RELU)
        if head is not None:
            outs = outs + head_out + head_outb
        return outs, dict(acts=acts, actTs=actTs, Wb=Wb, outs=outs if final == "relu" else None,
                          head=(head["kind"], head_saved) if head is not None else None)

    @staticmethod
    def _bwd_bf16(ctx, grads):
        G, NL, final, precision, masks, opts = ctx.cfg
        acts, actTs = ctx.saved["acts"], ctx.saved["actTs"]
        Ws = ctx.Ws
        dev = acts[0][0].device
        dWs = [[None] * NL for _ in range(G)]
        dbs = [[None] * NL for _ in range(G)]
        dxs = [None] * G
        dYb, dYT = [None] * G, [None] * G
        mn = [all(wgrad_mn_ok(*Ws[g][l].shape) for g in range(G)) for l in range(NL)]
        for g in range(G):
            M = acts[g][0].shape[0]
            N = Ws[g][-1].shape[0]
            dy = grads[g]
            if dy is None:
                dy = torch.zeros(M, N, dtype=torch.float32, device=dev)
            dy = _f32c(dy)
            if final == "relu":                      # the call ends in a ReLU: its mask is the saved output
                dy = dy * (ctx.saved["outs"][g] > 0)
            Np, Mp = (N + 7) // 8 * 8, (M + 7) // 8 * 8
            b = torch.empty(M, Np, dtype=torch.bfloat16, device=dev)
            bT = None if mn[NL - 1] else torch.empty(N, Mp, dtype=torch.bfloat16, device=dev)
            slot = _grad_slot(ctx.bs[g][NL - 1])
            db = slot if slot is not None else torch.zeros(N, dtype=torch.float32, device=dev)
            check(lib.dmf_cast_dual_bf16(ptr(dy), dy.stride(0), ptr(b), Np, ptr(bT), Mp, ptr(db), M, N, stream()))
            dYb[g], dYT[g] = (b[:, :N] if Np != N else b), bT
            dbs[g][NL - 1] = None if slot is not None else db
        for l in range(NL - 1, -1, -1):
            wdescs, ddescs, wdirect = [], [], []
            nextb, nextT = [None] * G, [None] * G
            next32 = [None] * G
            for g in range(G):
                X = acts[g][l]                      # bf16 [M,K]
                M, K = X.shape
                N = dYb[g].shape[1]
                Mp = (M + 7) // 8 * 8
                slot = _grad_slot(Ws[g][l]) if (N >= 512 and K >= 128) else None     # pair kernel can accumulate
                wdirect.append(slot is not None)
                dW = slot if slot is not None else torch.zeros(N, K, dtype=torch.float32, device=dev)
                if mn[l]:       # dW[n, k] = sum_m dY[m, n] X[m, k] straight from the row-major activations
                    wdescs.append(dict(A=dYb[g], lda=dYb[g].stride(0), B=X, ldb=X.stride(0), out_f32=dW, ldo_f32=K,
                                       M=N, N=K, K=M, split_k=-1 if slot is not None else 0, mn_major=1))
                else:
                    XT = actTs[g][l]
                    if XT is None:
                        XT = transpose_bf16(X)          # [K, Mp]
                    wdescs.append(dict(A=dYT[g], lda=dYT[g].stride(0), B=XT, ldb=XT.stride(0), out_f32=dW, ldo_f32=K,
                                       M=N, N=K, K=M, split_k=-1 if slot is not None else 0))
                dWs[g][l] = None if slot is not None else dW
                if l == 0 and not ctx.in_needs_grad[g] and ctx.extra_needs_grad[g]:
                    De = ctx.extra_cols[g]
                    WT = cast_transpose_bf16(Ws[g][l][:, K - De:])      # [De, Np]
                    dX32 = torch.empty(M, De, dtype=torch.float32, device=dev)
                    ddescs.append(dict(A=dYb[g], lda=dYb[g].stride(0), B=WT, ldb=WT.stride(0), out_f32=dX32,
                                       ldo_f32=De, M=M, N=De, K=N))
                    next32[g] = dX32
                elif l > 0 or ctx.in_needs_grad[g]:
                    WT = cast_transpose_bf16(Ws[g][l])          # [K, Np]
                    d = dict(A=dYb[g], lda=dYb[g].stride(0), B=WT, ldb=WT.stride(0), M=M, N=K, K=N)
                    if l > 0:
                        Kp = (K + 7) // 8 * 8
                        dXb = torch.empty(M, Kp, dtype=torch.bfloat16, device=dev)
                        dXT = None if mn[l - 1] else torch.empty(K, Mp, dtype=torch.bfloat16, device=dev)
                        d.update(out_bf16=dXb, ldo_bf16=Kp, mask=X, ldmask=X.stride(0))
                        if dXT is not None:
                            d.update(out_t=dXT, ldo_t=Mp)
                        nextb[g] = dXb[:, :K] if Kp != K else dXb
                        nextT[g] = dXT
                    else:
                        dX32 = torch.empty(M, K, dtype=torch.float32, device=dev)
                        d.update(out_f32=dX32, ldo_f32=K)
                        next32[g] = dX32
                    ddescs.append(d)
            if len(set(wdirect)) > 1:        # the accumulate mode needs every group of the launch on the pair kernel
                for g in range(G):
                    if wdirect[g]:
                        K_, N_ = wdescs[g]["N"], wdescs[g]["M"]
                        dW = torch.zeros(N_, K_, dtype=torch.float32, device=dev)
                        wdescs[g].update(out_f32=dW, split_k=0)
                        dWs[g][l] = dW
            gemm_tc(wdescs, L.EPI_NONE)
            if ddescs:
                gemm_tc(ddescs, L.EPI_RELU_MASK if l > 0 else L.EPI_NONE)
            if l == 0:
                dxs = next32
            else:
                for g in range(G):
                    K = nextb[g].shape[1]
                    slot = _grad_slot(ctx.bs[g][l - 1])
                    db = slot if slot is not None else torch.zeros(K, dtype=torch.float32, device=dev)
                    check(lib.dmf_colsum_bf16(ptr(nextb[g]), nextb[g].stride(0), nextb[g].shape[0], K, ptr(db), stream()))
                    dbs[g][l - 1] = None if slot is not None else db
                dYb, dYT = nextb, nextT
        return dxs, dWs, dbs

    @staticmethod
    def backward(ctx, *grads):
        G, NL, final, precision, masks, opts = ctx.cfg
        if precision == "bf16" and ctx.saved.get("head") is not None:
            # outputs were (X_g ..., head_g ..., bf16 copies): fold the head's backward into the gradient of X
            kind, hs = ctx.saved["head"]
            gx = []
            with _Prof("head_bwd"):
                for g in range(G):
                    gpre, gout = grads[g], grads[G + g]
                    if gout is None:
                        gx.append(gpre)
                        continue
                    gout = _f32c(gout)
                    R, D = gout.shape
                    dx = _f32c(gpre).clone() if gpre is not None else torch.empty(R, D, dtype=torch.float32, device=gout.device)
                    acc = 1 if gpre is not None else 0
                    if kind == 0:
                        y, inv = hs[g]
                        check(lib.dmf_row_normalize_bwd(ptr(y), D, ptr(inv), ptr(gout), D, R, D, ptr(dx), D, acc, stream()))
                    else:
                        e, w_, v_ = hs[g]
                        check(lib.dmf_vmf_bwd(ptr(e), D, ptr(w_), ptr(v_), ptr(gout), D, R, D, ptr(dx), D, acc, stream()))
                    gx.append(dx)
            grads = gx
        with _Prof("mlp_bwd"):
            if precision == "bf16":
                dxs, dWs, dbs = _GroupedMLP._bwd_bf16(ctx, grads)
            else:
                dxs, dWs, dbs = _GroupedMLP._bwd_f32(ctx, grads)
        out = [None]
        out += [dxs[g] if ctx.in_needs_grad[g] else None for g in range(G)]
        out += [dxs[g] if ctx.extra_needs_grad[g] else None for g in range(G)]
        for g in range(G):
            out += dWs[g]
        for g in range(G):
            out += dbs[g]
        return tuple(out)


def grouped_mlp(xs: Sequence[Tensor], weights: Sequence[Sequence[Tensor]], biases: Sequence[Sequence[Tensor]],
                final: str = "none", precision: str = "fp32", dropout_masks=None, extras=None, opts=None) -> List[Tensor]:
    """Apply G independent MLPs (Linear->ReLU)*(L-1)->Linear in L launches.  ``weights[g]`` /
    ``biases[g]`` list the L layers of group g (nn.Linear layout [out,in]).  ``extras[g]`` (optional)
    is appended column-wise to input g (torch.cat([x, extra], 1) semantics; fp32 path concatenates,
    bf16 path expects x to be the pre-cast bf16 concat buffer).  ``opts`` (bf16 path only): ``xTs`` = transposed
    bf16 copies [K0, Mp] of the inputs (K-major operand of the layer-0 wgrad), ``out_bf16`` / ``out_bf16T`` =
    per-group destinations (views into a later concat buffer) that also receive bf16 / transposed-bf16 copies
    of the final output from the last GEMM's epilogue, ``extras_prefilled`` = the tail columns of x already
    hold bf16(extra)."""
    G, NL = len(xs), len(weights[0])
    if precision == "bf16" and final == "evidence":
        # evidential heads in the bf16 path: the wide hidden layers (K = embedding width) run on the tensor cores
        # (bf16 operands, fp32 accumulate, ReLU in the epilogue); the narrow evidence layer (hidden -> C) and its
        # activation stay in fp32 -- exp() amplifies logit errors, and that layer is < 2 % of the head's FLOPs
        if NL < 2 or extras is not None:
            precision = "fp32"
        else:
            if dropout_masks is not None and NL > 2:
                raise NotImplementedError("bf16 evidential heads: dropout on inner hidden layers")
            hid = grouped_mlp(xs, [list(w[:-1]) for w in weights], [list(b[:-1]) for b in biases], final="relu",
                              precision="bf16")
            if dropout_masks is not None:
                hid = [h if (m is None or m[-1] is None) else h * m[-1] for h, m in zip(hid, dropout_masks)]
            return grouped_mlp(hid, [[w[-1]] for w in weights], [[b[-1]] for b in biases], final="evidence",
                               precision="fp32")
    if final == "relu" and precision != "bf16":
        raise L.DmfError("grouped_mlp: final='relu' is an internal mode of the bf16 path")
    flat = list(xs) + (list(extras) if extras is not None else [None] * G)
    for g in range(G):
        flat += list(weights[g])
    for g in range(G):
        flat += list(biases[g])
    return list(_GroupedMLP.apply((G, NL, final, precision, dropout_masks, opts), *flat))


# ----------------------------------------------------------------------------------------
# K2  InfoNCE (SupConLoss default path, models/losses.py:17-101)
# ----------------------------------------------------------------------------------------
class GatheredPair:
    """bf16 / fp32 copies of the two views of one critic call plus their all-gathered versions.  Creating it
    LAUNCHES the NCCL all-gathers asynchronously (on NCCL's stream); ``wait()`` joins them into the current stream.
    A module that knows all its critic inputs early (DisentangledSSL) creates the four pairs up front so that
    the gathers of calls 2..4 overlap the tiles of call 1."""

    def __init__(self, z0: Tensor, z1: Tensor, precision: str, bf16: Optional[Tuple[Tensor, Tensor]] = None):
        """``bf16`` = bf16 copies of (z0, z1) already written by the producing kernels (vMF / row-normalise epilogue)."""
        L.require_device()
        with _Prof("gather_launch"):
            z0, z1 = _f32c(z0.detach()), _f32c(z1.detach())
            if precision == "bf16" and bf16 is not None:
                self.a0, self.a1 = bf16
            else:
                self.a0, self.a1 = (cast_bf16(z0), cast_bf16(z1)) if precision == "bf16" else (z0, z1)
            self.works = []
            if _dist_on():
                world = dist.get_world_size()
                Bl, D = z0.shape
                self.g0 = torch.empty(Bl * world, D, dtype=self.a0.dtype, device=z0.device)
                self.g1 = torch.empty(Bl * world, D, dtype=self.a1.dtype, device=z0.device)
                self.works.append(dist.all_gather_into_tensor(self.g0, self.a0, async_op=True))
                self.works.append(dist.all_gather_into_tensor(self.g1, self.a1, async_op=True))
            else:
                self.g0, self.g1 = self.a0, self.a1

    def wait(self, also: Optional["torch.cuda.Stream"] = None):
        """make the current stream (and ``also``: the side stream of a ``_Fork``) wait for the gathers"""
        with _Prof("gather_wait"):
            for w in self.works:
                w.wait()
        if also is not None and self.works:
            with torch.cuda.stream(also):
                for w in self.works:
                    w.wait()
        self.works = []


_SIDE = {}
_FORK = os.environ.get("DMF_FORK", "1") != "0"
# Keep exp(s - shift) of every cross block (bf16, B_local x B bytes x 2 per critic call) between forward and backward so
# that the backward is one product per gradient instead of recompute + product (dmf_infonce_bwd_stored).  "0" restores the
# recompute kernels; DMF_STORE_E_MAX_GB bounds the bytes one batched op may keep (beyond it the calls fall back one by one).
_STORE_E = os.environ.get("DMF_STORE_E", "1") != "0"
_STORE_E_MAX = float(os.environ.get("DMF_STORE_E_MAX_GB", "64")) * 2 ** 30
# data parallel, stored probabilities: gradient of the view-1 rows as partial sums over this rank's E rows + one NCCL
# reduce-scatter per critic call ("0": recompute that direction from the gathered anchors with the M = 128 pair kernel)
_DP_RS = os.environ.get("DMF_DP_RS", "1") != "0"


class _Fork:
    """Deal INDEPENDENT kernel launches round-robin over the current stream and one side stream: the tail wave and the
    prologue of consecutive launches then overlap instead of serialising on stream order (each InfoNCE launch fills the
    GPU with one CTA per SM, so nothing else is shared).  Inside a CUDA-graph capture the fork / join become graph edges.
    Nothing may be allocated inside ``with fork.next():`` (the caching allocator is per stream)."""

    def __init__(self, dev):
        self.cur = torch.cuda.current_stream(dev)
        self.side = None
        if _FORK:
            key = (dev.index if dev.index is not None else torch.cuda.current_device())
            if key not in _SIDE:
                _SIDE[key] = torch.cuda.Stream(device=dev)
            self.side = _SIDE[key]
            self.side.wait_stream(self.cur)
        self.k = 0

    def resync(self):
        """the side stream must also see what the current stream has waited for since the fork (NCCL gathers)"""
        if self.side is not None:
            self.side.wait_stream(self.cur)

    def next(self):
        use_side = self.side is not None and (self.k & 1) == 1
        self.k += 1
        return torch.cuda.stream(self.side if use_side else self.cur)

    def join(self):
        if self.side is not None:
            self.cur.wait_stream(self.side)


class _InfoNCE(torch.autograd.Function):
    """forward(cfg, z0_0, z1_0, z0_1, z1_1, ...) -> out [ncalls, 3] = (loss, loss_x, loss_y) per critic call.
    cfg = (temperature, precision, bound, pres, reduce, diag_flags).  All calls of a step go through ONE invocation so
    that, under data parallelism, the column sums of every call are all-reduced by ONE collective and the row LSEs of
    every call are all-gathered by ONE collective (the kernels of all calls run first)."""

    @staticmethod
    def forward(ctx, cfg, *zs):
        L.require_device()
        temperature, precision, bound, pres, reduce, diag_flags = cfg[:6]
        layout = cfg[6] if len(cfg) > 6 else None
        zs = [_f32c(z) for z in zs]
        ctx.layout = layout
        if layout is not None:
            # zs are ROW-STACKED tensors; call c contrasts rows [ra, ra + Bl) of stack ia with rows [rb, rb + Bl) of stack
            # ib.  Slicing here, inside the Function, keeps autograd from embedding every half into a zero-filled
            # full-size gradient and adding the halves up again (5 extra passes over each stacked tensor per step).
            Bl = layout[0][4]
            ctx.stack_shapes = [tuple(z.shape) for z in zs]
            zs = [t for (ia, ra, ib, rb, _) in layout for t in (zs[ia][ra:ra + Bl], zs[ib][rb:rb + Bl])]
        nc = len(zs) // 2
        Bl, D = zs[0].shape
        if any(tuple(z.shape) != (Bl, D) for z in zs):
            raise L.DmfError("infonce_multi: every critic input of one batched op must have the same [rows, width]")
        dev = zs[0].device
        world = dist.get_world_size() if _dist_on() else 1
        rank = dist.get_rank() if world > 1 else 0
        Bg = Bl * world
        dt = 1 if precision == "bf16" else 0
        if pres is None:      # embeddings all-gathered over NVLink so every rank sees global negatives
            pres = [GatheredPair(zs[2 * c], zs[2 * c + 1], precision) for c in range(nc)]
        scale = 1.0 / temperature
        off = rank * Bl
        out3 = torch.zeros(nc, 3, dtype=torch.float32, device=dev)
        lse = torch.empty(nc, 2, Bl, dtype=torch.float32, device=dev)
        # the choice must be the SAME on every rank (the two paths issue different collective sequences): it may
        # depend on the shard size, never on this rank's row offset (off = rank * Bl is a multiple of 256 iff Bl is)
        fused = bound is not None and dt == 1 and D % 64 == 0 and D <= 512 and (world == 1 or Bl % 256 == 0)
        if fused:
            # unit-norm embeddings: fixed shift, row AND column sums from one pass over S01, symmetric
            # intra-view blocks on a half window (two B x B blocks of work instead of four)
            shift = float(bound)
            nrow = [3 if d else 1 for d in diag_flags]
            rbase = [sum(nrow[:c]) for c in range(nc)]
            rs = torch.zeros(sum(nrow), Bl, dtype=torch.float32, device=dev)
            cs = torch.zeros(sum(nrow), Bg, dtype=torch.float32, device=dev)
            dg = torch.empty(sum(nrow), Bl, dtype=torch.float32, device=dev)

            # stored probabilities (world == 1: both gradient directions; data parallel: the local-row direction)
            ebytes = int(lib.dmf_infonce_e_bytes(Bl, Bg)) if (_STORE_E and D in (256, 512)
                                                              and any(ctx.needs_input_grad[1:])) else 0
            estore = [None] * nc
            if ebytes:
                for c in range(min(nc, int(_STORE_E_MAX // ebytes))):
                    estore[c] = torch.empty(ebytes, dtype=torch.uint8, device=dev)

            def rowcol(A, Bm, k, sym, E=None):
                check(lib.dmf_infonce_rowcol_sums_store(ptr(A), A.stride(0), Bl, ptr(Bm), Bm.stride(0), Bg, D, scale, shift,
                                                        sym, off, ptr(rs[k]), ptr(cs[k]), off, ptr(dg[k]), ptr(E), stream()))
            # the launches write disjoint rows of rs / cs / dg: dealt over two streams so that their tails overlap; the
            # gathers of call c are awaited (by both streams) right before its launches, so later gathers stay in flight
            # under the tiles of the earlier calls
            fork = _Fork(dev)
            for c in range(nc):
                pres[c].wait(also=fork.side)
                a0, a1, g0, g1 = pres[c].a0, pres[c].a1, pres[c].g0, pres[c].g1
                b = rbase[c]
                with _Prof("rowlse_x4"):
                    with fork.next():
                        rowcol(a0, g1, b, 0, estore[c])   # cross block: rows -> view-0 anchors, columns -> view-1 anchors
                    if diag_flags[c]:
                        with fork.next():
                            rowcol(a0, g0, b + 1, 1)     # intra-view blocks: only the no-grad diagnostics loss_x / loss_y need them
                        with fork.next():
                            rowcol(a1, g1, b + 2, 1)     # (the reference's row max is the self-similarity 1/T = the fixed shift)
                    if c == nc - 1:
                        fork.join()
            if world > 1:
                with _Prof("colsum_allreduce"):
                    dist.all_reduce(cs)              # ONE collective for the column sums of every call
            mfix = torch.full((Bl,), shift, dtype=torch.float32, device=dev)
            with _Prof("finalize"):
                for c in range(nc):
                    b = rbase[c]
                    l_c1 = cs[b, off:off + Bl].contiguous()
                    if diag_flags[c]:
                        l_i0 = rs[b + 1] + cs[b + 1, off:off + Bl]
                        l_i1 = rs[b + 2] + cs[b + 2, off:off + Bl]
                        d1, d2 = dg[b + 1], dg[b + 2]
                    else:                            # callers that discard loss_x / loss_y (out3[c, 1:] is then meaningless)
                        l_i0, l_i1 = rs[b], l_c1
                        d1 = d2 = dg[b]
                    check(lib.dmf_infonce_finalize(ptr(mfix), ptr(rs[b]), ptr(mfix), ptr(l_i0), ptr(dg[b]), ptr(d1), Bl,
                                                   1.0 / (2 * Bg), 1.0 / Bg, 0, ptr(lse[c, 0]), ptr(out3[c]), stream()))
                    check(lib.dmf_infonce_finalize(ptr(mfix), ptr(l_c1), ptr(mfix), ptr(l_i1), ptr(dg[b]), ptr(d2), Bl,
                                                   1.0 / (2 * Bg), 1.0 / Bg, 1, ptr(lse[c, 1]), ptr(out3[c]), stream()))
        else:
            st = torch.empty(12, Bl, dtype=torch.float32, device=dev)
            wsb = lib.dmf_rowlse_workspace_bytes(Bl, Bg) if dt == 1 else 0
            ws = torch.empty(max(1, wsb // 4), dtype=torch.float32, device=dev)

            def rowlse(A, Bm, mo, lo, do):
                check(lib.dmf_rowlse(ptr(A), A.stride(0), Bl, ptr(Bm), Bm.stride(0), Bg, D, scale, ptr(st[mo]), ptr(st[lo]),
                                     off, ptr(st[do]), ptr(ws), wsb, dt, stream()))
            for c in range(nc):
                pres[c].wait()
                a0, a1, g0, g1 = pres[c].a0, pres[c].a1, pres[c].g0, pres[c].g1
                with _Prof("rowlse_x4"):
                    rowlse(a0, g1, 0, 1, 2)      # anchors z0 vs all z1: cross block, diag = positive
                    rowlse(a0, g0, 3, 4, 5)      # anchors z0 vs all z0: intra-view block, diag = self similarity
                    rowlse(a1, g0, 6, 7, 8)
                    rowlse(a1, g1, 9, 10, 11)
                check(lib.dmf_infonce_finalize(ptr(st[0]), ptr(st[1]), ptr(st[3]), ptr(st[4]), ptr(st[2]), ptr(st[5]), Bl,
                                               1.0 / (2 * Bg), 1.0 / Bg, 0, ptr(lse[c, 0]), ptr(out3[c]), stream()))
                check(lib.dmf_infonce_finalize(ptr(st[6]), ptr(st[7]), ptr(st[9]), ptr(st[10]), ptr(st[8]), ptr(st[11]), Bl,
                                               1.0 / (2 * Bg), 1.0 / Bg, 1, ptr(lse[c, 1]), ptr(out3[c]), stream()))
        if world > 1:
            with _Prof("lse_gather"):
                if reduce:
                    dist.all_reduce(out3)
                gath = torch.empty(world, nc, 2, Bl, dtype=torch.float32, device=dev)   # one collective: all calls, both views
                dist.all_gather_into_tensor(gath, lse)
                lse_all = gath.permute(1, 2, 0, 3).reshape(nc, 2, Bg).contiguous()
        else:
            lse_all = lse
        saved = []
        for c in range(nc):
            saved += [pres[c].a0, pres[c].a1, pres[c].g0, pres[c].g1]
        ctx.save_for_backward(lse, lse_all, *saved)
        ctx.meta = (nc, Bl, Bg, D, scale, off, dt)
        # not through save_for_backward: these are private byte buffers no hook / version counter needs to see
        ctx.estore = estore if fused else [None] * nc
        ctx.shift = float(bound) if fused else 0.0
        ctx.world = world
        return out3

    @staticmethod
    def backward(ctx, gout):
        lse, lse_all, *saved = ctx.saved_tensors
        nc, Bl, Bg, D, scale, off, dt = ctx.meta
        dev = lse.device
        gout = gout.contiguous().float()
        coef = scale / (2.0 * Bg)
        grads = []
        work = []
        layout = ctx.layout
        if layout is not None:
            # gradients of the row-stacked inputs: every critic call writes its two row blocks in place
            covered = [0] * len(ctx.stack_shapes)
            for (ia, ra, ib, rb, bl) in layout:
                covered[ia] += bl
                covered[ib] += bl
            sgr = [(torch.empty if covered[i] == shp[0] else torch.zeros)(shp, dtype=torch.float32, device=dev)
                   for i, shp in enumerate(ctx.stack_shapes)]
        for c in range(nc):
            a0, a1, g0, g1 = saved[4 * c: 4 * c + 4]
            gs = gout[c, 0:1]                           # only the loss has a gradient; diagnostics are no-grad
            if layout is not None:
                ia, ra, ib, rb, _ = layout[c]
                dz0, dz1 = sgr[ia][ra:ra + Bl], sgr[ib][rb:rb + Bl]
            else:
                dz0 = torch.empty(Bl, D, dtype=torch.float32, device=dev)
                dz1 = torch.empty(Bl, D, dtype=torch.float32, device=dev)
            if dt == 1 and lib.dmf_infonce_bwd_needs_transposed(D):
                with _Prof("transpose_gathered"):
                    g0T, g1T = transpose_bf16(g0), transpose_bf16(g1)
            else:       # fp32 path, or the bf16 kernel that reads the column block MN-major (D = 256 / 512)
                g0T = g1T = None
            E = ctx.estore[c]
            if E is not None:
                # E holds exp(s - shift) of rows = this rank's view-0 anchors, columns = all view-1 rows
                wk = torch.empty(int(lib.dmf_infonce_bwd_stored_work_floats(Bl, Bg)), dtype=torch.float32, device=dev)
                work.append(("E", E, lse[c, 0], lse_all[c, 1], g1, 0, gs, dz0, wk, None))
                wk1 = torch.empty_like(wk)
                if ctx.world == 1:
                    work.append(("E", E, lse[c, 0], lse_all[c, 1], a0, 1, gs, dz1, wk1, None))
                elif _DP_RS:
                    # data parallel: this rank's rows of E give a PARTIAL gradient for every column j (all ranks' view-1
                    # rows); one reduce-scatter per call sums the partials and leaves each rank its own rows.  The
                    # positive term -2 z0[j] is added by the rank that owns row j (diag_offset = its row offset).
                    full = torch.empty(Bg, D, dtype=torch.float32, device=dev)
                    work.append(("E", E, lse[c, 0], lse_all[c, 1], a0, 1, gs, full, wk1, dz1))
                else:       # DMF_DP_RS=0: recompute this direction from the gathered anchors instead
                    work.append(("R", a1, lse[c, 1], g0, g0T, lse_all[c, 0], gs, dz1))
            else:
                work.append(("R", a0, lse[c, 0], g1, g1T, lse_all[c, 1], gs, dz0))
                work.append(("R", a1, lse[c, 1], g0, g0T, lse_all[c, 0], gs, dz1))
            grads += [dz0, dz1]
        # 2 * nc independent launches (every one writes its own dz): dealt over two streams, tails overlap
        fork = _Fork(dev)
        pending = []
        with _Prof("infonce_bwd"):
            for w in work:
                with fork.next():
                    if w[0] == "E":
                        _, E, la, lb, z, direction, gs, dz, wk, scatter_to = w
                        check(lib.dmf_infonce_bwd_stored(ptr(E), Bl, Bg, ptr(la), ptr(lb), ctx.shift, ptr(z), z.stride(0), D,
                                                         direction, coef, ptr(gs), off, ptr(dz), D, 0, ptr(wk), stream()))
                        if scatter_to is not None:      # NCCL's stream picks up after this launch; later launches overlap it
                            pending.append(dist.reduce_scatter_tensor(scatter_to, dz, async_op=True))
                    else:
                        _, a, la, g, gT, lb, gs, dz = w
                        check(lib.dmf_infonce_bwd(ptr(a), a.stride(0), Bl, ptr(la), ptr(g), g.stride(0), ptr(gT),
                                                  gT.stride(0) if gT is not None else 0, Bg, ptr(lb), D, scale, coef, ptr(gs),
                                                  off, ptr(dz), D, 0, dt, stream()))
            fork.join()
        if pending:
            with _Prof("grad_reduce_scatter"):
                for h in pending:
                    h.wait()
        ctx.estore = [None] * nc        # free the probabilities now; a second backward (retain_graph) recomputes instead
        if layout is not None:
            return (None, *sgr)
        return (None, *grads)


def infonce_multi(pairs: Sequence[Tuple[Tensor, Tensor]], temperature: float = 0.07, precision: str = "fp32",
                  unit_norm: bool = False, pres: Optional[Sequence["GatheredPair"]] = None, reduce: bool = True,
                  diagnostics: Optional[Sequence[bool]] = None) -> Tensor:
    """All critic calls of a step in ONE op: returns out [ncalls, 3] = (loss, loss_x, loss_y) per call (rows of calls
    with ``diagnostics[c] == False`` carry the loss only).  Same semantics per call as ``infonce``; under
    torch.distributed the column sums / row LSEs / loss scalars of all calls share one collective each."""
    bound = (1.0 / float(temperature)) if unit_norm else None
    flags = tuple(bool(d) for d in (diagnostics if diagnostics is not None else [True] * len(pairs)))
    flat = [z for p in pairs for z in p]
    return _InfoNCE.apply((float(temperature), precision, bound, list(pres) if pres is not None else None, reduce, flags), *flat)


def infonce_stacked(stacks: Sequence[Tensor], layout: Sequence[Tuple[int, int, int, int]], rows: int,
                    temperature: float = 0.07, precision: str = "fp32", unit_norm: bool = False,
                    pres: Optional[Sequence["GatheredPair"]] = None, reduce: bool = True,
                    diagnostics: Optional[Sequence[bool]] = None) -> Tensor:
    """``infonce_multi`` on ROW-STACKED inputs: critic call c contrasts rows [ra, ra + rows) of ``stacks[ia]`` with rows
    [rb, rb + rows) of ``stacks[ib]`` for ``layout[c] = (ia, ra, ib, rb)``.  Same values and gradients as slicing the
    stacks and calling ``infonce_multi`` -- but the row blocks never pass through autograd as separate tensors, so the
    backward writes each block of d stack in place (no zero-filled embeddings, no adds)."""
    bound = (1.0 / float(temperature)) if unit_norm else None
    flags = tuple(bool(d) for d in (diagnostics if diagnostics is not None else [True] * len(layout)))
    lay = tuple((int(ia), int(ra), int(ib), int(rb), int(rows)) for ia, ra, ib, rb in layout)
    return _InfoNCE.apply((float(temperature), precision, bound, list(pres) if pres is not None else None, reduce, flags, lay),
                          *stacks)


def infonce(z0: Tensor, z1: Tensor, temperature: float = 0.07, precision: str = "fp32",
            unit_norm: bool = False, pre: Optional["GatheredPair"] = None, reduce: bool = True,
            diagnostics: bool = True) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """(loss, loss_x, loss_y) of SupConLoss()(stack([z0,z1],1)) without the [2B,2B] logits.
    Under torch.distributed the negatives are global (embeddings all-gathered with NCCL).
    ``unit_norm=True`` asserts that the rows of z0 / z1 are L2-normalised (|s| <= 1/T): the bf16 path then runs
    the fixed-shift row+column kernel (dmf_infonce_rowcol_sums) instead of four online-max passes.
    ``pre`` = a GatheredPair built earlier from the same (z0, z1) (asynchronous all-gather already in flight).
    ``reduce=False`` (data parallel only) returns this rank's PARTIAL sums of the three scalars -- the caller
    all-reduces them later, once for all its critic calls (the gradients do not depend on the loss value).
    ``diagnostics=False``: the caller discards loss_x / loss_y (the reference's specific-critic calls,
    models/disentangledssl.py:143-144); the unit-norm bf16 path then skips the two intra-view blocks (half of the
    forward work) and returns None for them."""
    out = infonce_multi([(z0, z1)], temperature, precision, unit_norm, [pre] if pre is not None else None, reduce,
                        [diagnostics])[0]
    if not diagnostics:
        return out[0], None, None
    return out[0], out[1].detach(), out[2].detach()


# ----------------------------------------------------------------------------------------
# row normalisation, vMF sample, ortho loss
# ----------------------------------------------------------------------------------------
class _RowNormalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, want_bf16=False):
        L.require_device()
        x = _f32c(x)
        R, D = x.shape
        y = torch.empty_like(x)
        yb = torch.empty(R, D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        inv = torch.empty(R, dtype=torch.float32, device=x.device)
        with _Prof("head_fwd"):
            check(lib.dmf_row_normalize_fwd(ptr(x), D, R, D, eps, ptr(y), D, ptr(yb), D, ptr(inv), stream()))
        ctx.save_for_backward(y, inv)
        if want_bf16:
            ctx.mark_non_differentiable(yb)
            return y, yb
        return y

    @staticmethod
    def backward(ctx, dy, *unused):
        y, inv = ctx.saved_tensors
        dy = _f32c(dy)
        R, D = y.shape
        dx = torch.empty_like(y)
        with _Prof("head_bwd"):
            check(lib.dmf_row_normalize_bwd(ptr(y), D, ptr(inv), ptr(dy), D, R, D, ptr(dx), D, 0, stream()))
        return dx, None, None


def row_normalize(x: Tensor, eps: float = 1e-12, want_bf16: bool = False):
    """F.normalize(x, dim=-1); ``want_bf16`` also returns a bf16 copy written by the same kernel pass."""
    return _RowNormalize.apply(x, eps, want_bf16)


class _VmfSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, w, v, want_bf16=False):
        L.require_device()
        e, w, v = _f32c(e), _f32c(w), _f32c(v)
        R, D = e.shape
        z = torch.empty_like(e)
        zb = torch.empty(R, D, dtype=torch.bfloat16, device=e.device) if want_bf16 else None
        with _Prof("head_fwd"):
            check(lib.dmf_vmf_fwd(ptr(e), D, ptr(w), ptr(v), R, D, ptr(z), D, ptr(zb), D, stream()))
        ctx.save_for_backward(e, w, v)
        if want_bf16:
            ctx.mark_non_differentiable(zb)
            return z, zb
        return z

    @staticmethod
    def backward(ctx, dz, *unused):
        e, w, v = ctx.saved_tensors
        dz = _f32c(dz)
        R, D = e.shape
        de = torch.empty_like(e)
        with _Prof("head_bwd"):
            check(lib.dmf_vmf_bwd(ptr(e), D, ptr(w), ptr(v), ptr(dz), D, R, D, ptr(de), D, 0, stream()))
        return de, None, None, None


def vmf_rsample(e: Tensor, w: Tensor, v: Tensor, want_bf16: bool = False):
    """ProbabilisticEncoder('vmf')(e)[0].rsample() with the noise (w [B,1], v [B,D-1]) supplied; ``want_bf16`` also
    returns a bf16 copy of the sample written by the same kernel pass."""
    return _VmfSample.apply(e, w, v, want_bf16)


class _OrthoLoss(torch.autograd.Function):
    """|| normalize(z1)^T normalize(zs) ||_F  (models/losses.py:104-110); Gram partial sums are
    all-reduced across ranks before the square root."""

    @staticmethod
    def forward(ctx, z1, zs, precision):
        L.require_device()
        z1, zs = _f32c(z1), _f32c(zs)
        R, D = z1.shape
        dev = z1.device
        n1, ns = torch.empty_like(z1), torch.empty_like(zs)
        i1 = torch.empty(R, dtype=torch.float32, device=dev)
        i2 = torch.empty(R, dtype=torch.float32, device=dev)
        check(lib.dmf_row_normalize_fwd(ptr(z1), D, R, D, 1e-12, ptr(n1), D, 0, 0, ptr(i1), stream()))
        check(lib.dmf_row_normalize_fwd(ptr(zs), D, R, D, 1e-12, ptr(ns), D, 0, 0, ptr(i2), stream()))
        # split the batch (the contraction dim) over chunks so the grid fills the GPU
        if precision == "bf16" and R >= 64:
            # tensor-core Gram: both operands K-major after a cast+transpose ([D, R] bf16)
            n1T, nsT = cast_transpose_bf16(n1), cast_transpose_bf16(ns)
            Rp = n1T.stride(0)
            tiles = ((D + 127) // 128) ** 2
            chunks = max(1, min(48, (2 * 148 + tiles - 1) // tiles, R // 256))
            rows = ((R + chunks - 1) // chunks + 63) // 64 * 64
            chunks = (R + rows - 1) // rows
            part = torch.empty(chunks, D * D, dtype=torch.float32, device=dev)
            descs = []
            for c in range(chunks):
                r0 = c * rows
                rc = min(rows, R - r0)
                descs.append(dict(A=n1T[:, r0:], lda=Rp, B=nsT[:, r0:], ldb=Rp, out_f32=part[c], ldo_f32=D, M=D, N=D, K=rc))
            gemm_tc(descs, L.EPI_NONE)
        else:
            chunks = max(1, min(64, R // 512))
            rows = (R + chunks - 1) // chunks
            chunks = (R + rows - 1) // rows
            part = torch.empty(chunks, D * D, dtype=torch.float32, device=dev)
            descs = []
            for c in range(chunks):
                r0 = c * rows
                rc = min(rows, R - r0)
                descs.append(dict(A=n1[r0:], a_rs=1, a_cs=D, B=ns[r0:], b_rs=D, b_cs=1, C=part[c], ldc=D, M=D, N=D, K=rc))
            gemm_f32(descs, L.EPI_NONE)
        gram = colsum(part).view(D, D)
        if _dist_on():
            dist.all_reduce(gram)
        ss = torch.zeros(1, dtype=torch.float32, device=dev)
        check(lib.dmf_sumsq_f32(ptr(gram), D * D, ptr(ss), stream()))
        loss = torch.sqrt(ss)[0]
        ctx.save_for_backward(n1, ns, i1, i2, gram, loss)
        return loss

    @staticmethod
    def backward(ctx, g):
        n1, ns, i1, i2, gram, loss = ctx.saved_tensors
        R, D = n1.shape
        dev = n1.device
        dn1 = torch.empty_like(n1)
        dns = torch.empty_like(ns)
        # dN1 = Ns * M^T / L ; dNs = N1 * M / L
        gemm_f32([dict(A=ns, a_rs=D, a_cs=1, B=gram, b_rs=1, b_cs=D, C=dn1, ldc=D, M=R, N=D, K=D),
                  dict(A=n1, a_rs=D, a_cs=1, B=gram, b_rs=D, b_cs=1, C=dns, ldc=D, M=R, N=D, K=D)], L.EPI_NONE)
        sc = (g / loss).reshape(1, 1)
        dn1.mul_(sc)
        dns.mul_(sc)
        d1 = torch.empty_like(n1)
        d2 = torch.empty_like(ns)
        check(lib.dmf_row_normalize_bwd(ptr(n1), D, ptr(i1), ptr(dn1), D, R, D, ptr(d1), D, 0, stream()))
        check(lib.dmf_row_normalize_bwd(ptr(ns), D, ptr(i2), ptr(dns), D, R, D, ptr(d2), D, 0, stream()))
        return d1, d2, None


@torch.no_grad()
def ortho_values_nograd(pairs: Sequence[Tuple[Tensor, Tensor]], precision: str = "fp32") -> Tensor:
    """Values of ``ortho_loss`` for several (z1, zs) pairs whose rows are ALREADY L2-normalised (fp32, or the bf16
    copies the head kernels wrote), without autograd:
    one grouped split-K Gram launch (bf16 path) for all pairs, one all-reduce, one reduction.  Used by
    DisentangledSSL when the ortho weight is exactly zero (the term is only logged)."""
    L.require_device()
    n = len(pairs)
    R, D = pairs[0][0].shape
    dev = pairs[0][0].device
    grams = torch.zeros(n, D, D, dtype=torch.float32, device=dev)
    if precision == "bf16" and R >= 512 and D >= 128:
        Rp = (R + 7) // 8 * 8
        descs = []
        mn = wgrad_mn_ok(D, D)      # the Gram reads row-major bf16 rows MN-major: no transposed copies
        for i, (a, b) in enumerate(pairs):
            if mn:
                ab = a if a.dtype == torch.bfloat16 else cast_bf16(_f32c(a))
                bb = b if b.dtype == torch.bfloat16 else cast_bf16(_f32c(b))
                descs.append(dict(A=ab, lda=ab.stride(0), B=bb, ldb=bb.stride(0), out_f32=grams[i], ldo_f32=D, M=D, N=D,
                                  K=R, split_k=0, mn_major=1))
                continue
            a, b = a.float(), b.float()
            aT = torch.empty(D, Rp, dtype=torch.bfloat16, device=dev)
            bT = torch.empty(D, Rp, dtype=torch.bfloat16, device=dev)
            cast_dual_bf16(a, None, 0, aT, Rp)
            cast_dual_bf16(b, None, 0, bT, Rp)
            descs.append(dict(A=aT, lda=Rp, B=bT, ldb=Rp, out_f32=grams[i], ldo_f32=D, M=D, N=D, K=R, split_k=0))
        gemm_tc(descs, L.EPI_NONE)
    else:
        for i, (a, b) in enumerate(pairs):
            a, b = _f32c(a.float()), _f32c(b.float())
            chunks = max(1, min(64, R // 512))
            rows = (R + chunks - 1) // chunks
            chunks = (R + rows - 1) // rows
            part = torch.empty(chunks, D * D, dtype=torch.float32, device=dev)
            descs = []
            for c in range(chunks):
                r0 = c * rows
                rc = min(rows, R - r0)
                descs.append(dict(A=a[r0:], a_rs=1, a_cs=D, B=b[r0:], b_rs=D, b_cs=1, C=part[c], ldc=D, M=D, N=D, K=rc))
            gemm_f32(descs, L.EPI_NONE)
            colsum(part, grams[i].view(-1))
    if _dist_on():
        dist.all_reduce(grams)
    return torch.sqrt((grams * grams).sum(dim=(1, 2)))


def ortho_loss(z1: Tensor, zs: Tensor, precision: str = "fp32") -> Tensor:
    return _OrthoLoss.apply(z1, zs, precision)


# ----------------------------------------------------------------------------------------
# DMVAE head + reconstruction  (models/dmvae.py:74-176)
# ----------------------------------------------------------------------------------------
class _DmvaeHead(torch.autograd.Function):
    """forward(cfg, noise, *stats) -> (dec_in_0..dec_in_{N-1}, kl3).  cfg = (N, B, e, T).
    kl3 = (kl_private, kl_poe, kl_uni) batch means; differentiable through the same backward kernel."""

    @staticmethod
    def forward(ctx, cfg, noise, *stats):
        L.require_device()
        N, B, e, T = cfg
        stats = [_f32c(s) for s in stats]
        noise = _f32c(noise)
        dev = stats[0].device
        dec_in = [torch.empty(N * B, 2 * e, dtype=torch.float32, device=dev) for _ in range(N)]
        kl3 = torch.zeros(3, dtype=torch.float32, device=dev)
        check(lib.dmf_dmvae_head_fwd(L.ptr_array(stats), ptr(noise), N, B, e, T, L.ptr_array(dec_in), ptr(kl3), stream()))
        ctx.cfg = cfg
        ctx.save_for_backward(noise, *stats)
        return (*dec_in, kl3)

    @staticmethod
    def backward(ctx, *grads):
        N, B, e, T = ctx.cfg
        noise, *stats = ctx.saved_tensors
        dev = noise.device
        dd = [(_f32c(g) if g is not None else torch.zeros(N * B, 2 * e, dtype=torch.float32, device=dev))
              for g in grads[:N]]
        gk = grads[N]
        gk = _f32c(gk) if gk is not None else torch.zeros(3, dtype=torch.float32, device=dev)
        dst = [torch.empty(B, 4 * e, dtype=torch.float32, device=dev) for _ in range(N)]
        check(lib.dmf_dmvae_head_bwd(L.ptr_array(stats), ptr(noise), L.ptr_array(dd), N, B, e, T, ptr(gk),
                                     L.ptr_array(dst), stream()))
        return (None, None, *dst)


class _DmvaeMse(torch.autograd.Function):
    """forward(cfg, recon, x) -> out2 = (w_joint*mse_joint, w_cross*sum mse_cross); the gradient
    w.r.t. recon is produced by the same kernel pass.  cfg = (N, B, d, view, w_joint, w_cross)."""

    @staticmethod
    def forward(ctx, cfg, recon, x):
        L.require_device()
        N, B, d, view, wj, wc = cfg
        recon, x = _f32c(recon), _f32c(x)
        out2 = torch.zeros(2, dtype=torch.float32, device=recon.device)
        dr = torch.empty_like(recon)
        check(lib.dmf_dmvae_mse_fwd_bwd(ptr(recon), d, ptr(x), d, N, B, d, view, wj, wc, 0, ptr(out2), ptr(dr), d, stream()))
        ctx.save_for_backward(dr)
        return out2

    @staticmethod
    def backward(ctx, g):
        (dr,) = ctx.saved_tensors
        # both outputs enter the loss with the same upstream scale (loss = out2[0] + out2[1] + ...)
        return None, dr * g[0], None


# ----------------------------------------------------------------------------------------
# K3  fused evidence fusion + EDL loss  (utils.py:66-116, models/losses.py:117-248)
# ----------------------------------------------------------------------------------------
class _EdlLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, evid, labels, agg, coef, dc_weight, inv_B_global):
        L.require_device()
        evid = _f32c(evid)
        B, V, Cc = evid.shape
        dev = evid.device
        labels = labels.to(torch.int64).contiguous()
        fused = torch.empty(B, Cc, dtype=torch.float32, device=dev)
        grad = torch.empty_like(evid)
        parts = torch.zeros(4, dtype=torch.float32, device=dev)
        p = L.EdlParams(B, V, Cc, L.AGG[agg], coef, dc_weight, inv_B_global)
        with _Prof("edl_fused"):
            check(lib.dmf_edl_fused(ptr(evid), ptr(labels), p, 0, ptr(fused), ptr(grad), 0, 0, 0, ptr(parts), stream()))
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(fused)
        return fused, parts

    @staticmethod
    def backward(ctx, gf, gp):
        (grad,) = ctx.saved_tensors
        return grad * gp[3], None, None, None, None, None


def edl_fused_loss(evid: Tensor, labels: Tensor, agg: str, annealing_step: int, annealing_start: int,
                   fused: float = 1.0, gamma: float = 1.0, global_batch: Optional[int] = None):
    """AvgTrustedLoss.forward + the aggregation rule in one kernel pass.
    Returns (loss, fused_evidence [B,C], parts[4] = (edl, coef*kl, dc, total))."""
    B = evid.shape[0]
    coef = min(1.0, annealing_step / annealing_start)
    t = min(1.0, annealing_step / max(1, annealing_start))
    gamma_t = 0.2 * (1 - t) + gamma * t
    fe, parts = _EdlLoss.apply(evid, labels, agg, float(coef), float(gamma_t * fused),
                               1.0 / float(global_batch or B))
    return parts[3], fe, parts


def edl_summaries(evid: Tensor, labels: Tensor, agg: str):
    """Forward-only pass: fused evidence, u = C/S, aleatoric, per-view and fused argmax."""
    L.require_device()
    evid = _f32c(evid)
    B, V, Cc = evid.shape
    dev = evid.device
    fused = torch.empty(B, Cc, dtype=torch.float32, device=dev)
    u = torch.empty(B, dtype=torch.float32, device=dev)
    ale = torch.empty(B, dtype=torch.float32, device=dev)
    pred = torch.empty(B, V + 1, dtype=torch.int32, device=dev)
    p = L.EdlParams(B, V, Cc, L.AGG[agg], 0.0, 0.0, 1.0 / max(B, 1))
    labels = labels.to(torch.int64).contiguous()
    check(lib.dmf_edl_fused(ptr(evid), ptr(labels), p, 0, ptr(fused), 0, ptr(u), ptr(ale), ptr(pred), 0, stream()))
    return fused, u, ale, pred


def eval_reduce(evid: Tensor, fused: Tensor, labels: Tensor, acc: Optional[dict] = None) -> dict:
    """One-pass accumulation of the evaluation statistics of analysis.py:5-399 for a batch.  ``acc`` (returned by a
    previous call) keeps the device accumulators across batches: no per-batch ``.item()``."""
    L.require_device()
    evid, fused = _f32c(evid), _f32c(fused)
    B, V, Cc = evid.shape
    dev = evid.device
    if acc is None:
        acc = dict(stats=torch.zeros(V + 1, 8, dtype=torch.float32, device=dev),
                   class_sum=torch.zeros(V + 1, Cc, dtype=torch.float32, device=dev),
                   true_sum=torch.zeros(V + 1, Cc, dtype=torch.float32, device=dev),
                   class_counts=torch.zeros(Cc, dtype=torch.float32, device=dev), N=0)
    labels = labels.to(torch.int64).contiguous()
    check(lib.dmf_eval_reduce(ptr(evid), ptr(fused), ptr(labels), B, V, Cc, ptr(acc["stats"]), ptr(acc["class_sum"]),
                              ptr(acc["true_sum"]), ptr(acc["class_counts"]), stream()))
    acc["N"] += B
    return acc


def fuse_evidence(evid: Tensor, agg: str) -> Tensor:
    """utils.py:66-116 aggregation only (no gradient: the reference loss ignores it)."""
    B = evid.shape[0]
    dummy = torch.zeros(B, dtype=torch.int64, device=evid.device)
    L.require_device()
    evid = _f32c(evid)
    _, V, Cc = evid.shape
    fused = torch.empty(B, Cc, dtype=torch.float32, device=evid.device)
    p = L.EdlParams(B, V, Cc, L.AGG[agg], 0.0, 0.0, 1.0 / max(B, 1))
    check(lib.dmf_edl_fused(ptr(evid), ptr(dummy), p, 0, ptr(fused), 0, 0, 0, 0, 0, stream()))
    return fused


# ----------------------------------------------------------------------------------------
# optimizer (a18)
# ----------------------------------------------------------------------------------------
def adam_step_flat(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, step: int, betas=(0.9, 0.999),
                   eps: float = 1e-8, weight_decay: float = 0.0, decoupled: bool = False, grad_scale: float = 1.0) -> None:
    L.require_device()
    check(lib.dmf_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, betas[0], betas[1], eps, weight_decay,
                            1 if decoupled else 0, step, grad_scale, 0, stream()))


def adam_step_flat_dev(p: Tensor, g: Tensor, m: Tensor, v: Tensor, state: Tensor, betas=(0.9, 0.999), eps: float = 1e-8,
                       weight_decay: float = 0.0, decoupled: bool = False, grad_scale: float = 1.0) -> None:
    """Graph-capturable Adam/AdamW: ``state`` is a device float[2] = (steps taken so far, learning rate)."""
    L.require_device()
    check(lib.dmf_adam_step_dev(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(state), betas[0], betas[1], eps, weight_decay,
                                1 if decoupled else 0, grad_scale, 0, stream()))
