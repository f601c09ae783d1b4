"""H2D bandwidth of pinned host buffers with all ranks copying at once (torchrun); prints GB/s per rank."""
import os, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
res = []
for mb in (32, 64, 128):
    h = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    res.append((mb, mb * 10 / 1024 / (s.elapsed_time(e) / 1e3)))
aff = len(os.sched_getaffinity(0))
print(f"rank {rank}: " + ", ".join(f"{mb} MB: {g:.1f} GB/s" for mb, g in res) + f"  (cpu affinity {aff} cores)", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
