"""Times the collectives the data-parallel step uses (NCCL, one process per GPU); prints per-op ms and GB/s."""
import os
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bg, D = 65536, 512
    Bl = Bg // world

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / iters
    a = torch.randn(Bl, D, device=dev).bfloat16(); g = torch.empty(Bg, D, dtype=torch.bfloat16, device=dev)
    cs = torch.zeros(3, Bg, device=dev); gram = torch.zeros(D, D, device=dev); flat = torch.zeros(4_720_000, device=dev)
    lse = torch.zeros(Bl, device=dev); lse_all = torch.zeros(Bg, device=dev); o3 = torch.zeros(3, device=dev)
    res = {
        "all_gather emb bf16": (timeit(lambda: dist.all_gather_into_tensor(g, a)), g.numel() * 2),
        "all_reduce colsums [3,B]": (timeit(lambda: dist.all_reduce(cs)), cs.numel() * 4),
        "all_reduce gram [D,D]": (timeit(lambda: dist.all_reduce(gram)), gram.numel() * 4),
        "all_reduce flat grads": (timeit(lambda: dist.all_reduce(flat)), flat.numel() * 4),
        "all_gather lse": (timeit(lambda: dist.all_gather_into_tensor(lse_all, lse)), lse_all.numel() * 4),
        "all_reduce 3 scalars": (timeit(lambda: dist.all_reduce(o3)), 12),
    }
    if rank == 0:
        for k, (ms, b) in res.items():
            print(f"{k:28s} {ms*1e3:9.1f} us   {b/ms/1e6:8.1f} GB/s (payload/time)")
        print("NCCL", torch.cuda.nccl.version(), "P2P", torch.cuda.can_device_access_peer(0, 1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
