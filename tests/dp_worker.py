"""Worker for the multi-GPU data-parallel parity test (launched with torch.distributed.run, NCCL).

Every rank builds the same DisentangledSSL (same seed), takes its row shard of the same global batch and
explicit vMF noise, runs one forward/backward with global negatives (all-gathered embeddings) and
all-reduces the flat gradient.  Rank 0 then repeats the step on the FULL batch with the collectives switched
off and compares loss, logs and the summed gradient (bf16 path: both runs use the same kernels, so the
difference is only summation order)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import disentagled_multimodal_fusion_b200 as pkg
    from disentagled_multimodal_fusion_b200 import ops
    from disentagled_multimodal_fusion_b200.dp import FlatParams, shard_rows

    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    dims, h, e, Bg = [256, 192], 128, 128, 1024
    torch.manual_seed(0)
    model = pkg.DisentangledSSL(output_dim=dims, hidden_dim=h, embed_dim=e, precision=prec).to(dev)
    gen = torch.Generator().manual_seed(1)
    x1, x2 = torch.randn(Bg, dims[0], generator=gen).to(dev), torch.randn(Bg, dims[1], generator=gen).to(dev)
    v1 = x1 + 0.01 * torch.randn(Bg, dims[0], generator=gen).to(dev)
    v2 = x2 + 0.01 * torch.randn(Bg, dims[1], generator=gen).to(dev)
    torch.manual_seed(7)
    noise = model.draw_noise(Bg, dev)                       # reference-stream noise for the global batch
    lo, hi = shard_rows(Bg, rank, world)
    fp = FlatParams(model.parameters())

    fp.zero_grad()
    loss, logs = model(x1[lo:hi], x2[lo:hi], v1[lo:hi], v2[lo:hi], noise=[(w[lo:hi], v[lo:hi]) for w, v in noise])
    loss.backward()
    fp.allreduce_grads()
    g_dp = fp.grad.clone()
    torch.cuda.synchronize()

    ok = True
    if rank == 0:
        saved = ops._dist_on
        ops._dist_on = lambda: False
        try:
            fp.zero_grad()
            loss1, logs1 = model(x1, x2, v1, v2, noise=noise)
            loss1.backward()
        finally:
            ops._dist_on = saved
        g_1 = fp.grad
        tol = 2e-5 if prec == "fp32" else 2e-3

        def rel(a, b):
            return float((a - b).abs().max() / (b.abs().max() + 1e-30))
        r_loss = rel(loss.detach(), loss1.detach())
        r_grad = rel(g_dp, g_1)
        r_logs = max(rel(torch.as_tensor(logs[k], device=dev).float(), torch.as_tensor(logs1[k], device=dev).float())
                     for k in ("shared", "specific", "ortho"))
        print(f"dp{world} [{prec}] loss {float(loss):.6f} vs single {float(loss1):.6f}: rel {r_loss:.2e}; logs {r_logs:.2e}; "
              f"grad rel {r_grad:.2e}", flush=True)
        ok = r_loss < tol and r_grad < 10 * tol and r_logs < tol
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    print("OK", rank, flush=True)


if __name__ == "__main__":
    main()
