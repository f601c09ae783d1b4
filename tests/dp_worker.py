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
    import disentagled_multimodal_fusion_b200 as pkg
    from disentagled_multimodal_fusion_b200.dp import init_process_group_nccl
    init_process_group_nccl(dev)
    from disentagled_multimodal_fusion_b200 import ops, dp as dpmod
    from disentagled_multimodal_fusion_b200.dp import FlatParams, shard_rows

    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    case = sys.argv[2] if len(sys.argv) > 2 else "dssl"

    def rel(a, b):
        return float((a - b).abs().max() / (b.abs().max() + 1e-30))

    def single(fn):
        """run ``fn`` with the collectives switched off (the whole batch in this process)"""
        saved_ops, saved_dp = ops._dist_on, dpmod.world
        ops._dist_on = lambda: False
        dpmod.world = lambda: (0, 1)
        try:
            return fn()
        finally:
            ops._dist_on, dpmod.world = saved_ops, saved_dp

    ok = True
    if case in ("dssl", "dssl_small", "dssl_e256"):
        # dssl: per-rank shard 512 rows (fused row+column kernel); dssl_small: 128 rows per rank, where the fused
        # kernel is not eligible -- every rank must then take the generic path (rank-invariant choice)
        # dssl_e256: embed width 256 -> the stored-probability backward (local-row direction from this rank's E rows,
        # the other direction as partial sums + NCCL reduce-scatter)
        dims, h, e = [256, 192], 128, (256 if case == "dssl_e256" else 128)
        Bg = 128 * world if case == "dssl_small" else 1024
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
        if rank == 0:
            def full():
                fp.zero_grad()
                l1, lg1 = model(x1, x2, v1, v2, noise=noise)
                l1.backward()
                return l1, lg1
            loss1, logs1 = single(full)
            tol = 2e-5 if prec == "fp32" else 2e-3
            r_loss = rel(loss.detach(), loss1.detach())
            r_grad = rel(g_dp, fp.grad)
            r_logs = max(rel(torch.as_tensor(logs[k], device=dev).float(), torch.as_tensor(logs1[k], device=dev).float())
                         for k in ("shared", "specific", "ortho"))
            print(f"dp{world} {case} [{prec}] loss {float(loss):.6f} vs single {float(loss1):.6f}: rel {r_loss:.2e}; "
                  f"logs {r_logs:.2e}; grad rel {r_grad:.2e}", flush=True)
            ok = r_loss < tol and r_grad < 10 * tol and r_logs < tol
    elif case == "probe":
        # frozen DSSL backbone + evidential probe heads: EDL loss normalised by the GLOBAL batch, gradients summed
        dims, h, e, Bg, ncls = [64, 48], 64, 32, 512, 7
        torch.manual_seed(0)
        bbm = pkg.DisentangledSSL(output_dim=dims, hidden_dim=h, embed_dim=e, precision="fp32").to(dev)
        probe = pkg.EvidentialProbeModule(bbm, num_classes=ncls, input_dim=e, hidden_dim=(32,), dropout=0.0,
                                          annealing_start=10, aggregation="cml", fused=1).to(dev)
        probe.criterion.annealing_step = 4
        gen = torch.Generator().manual_seed(2)
        x1, x2 = torch.randn(Bg, dims[0], generator=gen).to(dev), torch.randn(Bg, dims[1], generator=gen).to(dev)
        y = torch.randint(0, ncls, (Bg,), generator=gen).to(dev)
        lo, hi = shard_rows(Bg, rank, world)
        fp = FlatParams([p for n, p in probe.named_parameters() if not n.startswith("backbone.")])
        fp.zero_grad()
        loss = probe.shared_step([x1[lo:hi], x2[lo:hi], y[lo:hi]])[0]
        loss.backward()
        fp.allreduce_grads()
        g_dp = fp.grad.clone()
        torch.cuda.synchronize()
        if rank == 0:
            def full():
                fp.zero_grad()
                l1 = probe.shared_step([x1, x2, y])[0]
                l1.backward()
                return l1
            loss1 = single(full)
            r_loss, r_grad = rel(loss.detach(), loss1.detach()), rel(g_dp, fp.grad)
            print(f"dp{world} probe loss {float(loss):.6f} vs single {float(loss1):.6f}: rel {r_loss:.2e}; grad rel {r_grad:.2e}",
                  flush=True)
            ok = r_loss < 2e-5 and r_grad < 2e-4
    elif case == "dmvae":
        dims, h, e, Bg = [40, 24, 16], 64, 12, 256
        torch.manual_seed(0)
        model = pkg.DMVAE(output_dim=dims, hidden_dim=h, embed_dim=e, a=1e-2).to(dev)
        gen = torch.Generator().manual_seed(3)
        xs = [torch.rand(Bg, d, generator=gen).to(dev) for d in dims]
        torch.manual_seed(11)
        noise = model.draw_noise(Bg, dev)
        lo, hi = shard_rows(Bg, rank, world)
        fp = FlatParams(model.parameters())
        fp.zero_grad()
        loss, logs = model([x[lo:hi] for x in xs], noise=noise[:, lo:hi].contiguous())
        loss.backward()
        fp.allreduce_grads()
        g_dp = fp.grad.clone()
        torch.cuda.synchronize()
        if rank == 0:
            def full():
                fp.zero_grad()
                l1, lg1 = model(xs, noise=noise)
                l1.backward()
                return l1, lg1
            loss1, logs1 = single(full)
            r_loss, r_grad = rel(loss.detach(), loss1.detach()), rel(g_dp, fp.grad)
            r_logs = max(rel(torch.as_tensor(logs[k], device=dev).float(), torch.as_tensor(logs1[k], device=dev).float())
                         for k in ("loss_joint_recon", "loss_cross_recon", "kl_private", "kl_shared_poe", "kl_shared_uni_sum"))
            print(f"dp{world} dmvae loss {float(loss):.6f} vs single {float(loss1):.6f}: rel {r_loss:.2e}; logs {r_logs:.2e}; "
                  f"grad rel {r_grad:.2e}", flush=True)
            ok = r_loss < 2e-5 and r_grad < 2e-4 and r_logs < 2e-5
    else:
        raise SystemExit(f"unknown case {case}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    print("OK", rank, flush=True)


if __name__ == "__main__":
    main()
