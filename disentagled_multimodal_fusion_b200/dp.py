"""Data-parallel host logic (one process per GPU, torch.distributed; NCCL on the box, gloo in the CPU
tests).  The path shards by batch rows: rank r owns rows [r*B/R, (r+1)*B/R) of every view, parameters
are replicated.  Exchange steps: embeddings + row-LSEs all-gathered inside
ops.infonce, the ortho Gram all-reduced inside ops.ortho_loss, and here the flat gradient
all-reduce (SUM: every loss term is already normalised by the GLOBAL batch)."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Row range [begin, end) owned by ``rank``; the global batch must divide evenly so that every
    rank contributes the same number of anchors to the all-gather."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def globalize_mean_losses(loss: torch.Tensor, logs: List[torch.Tensor]):
    """For a loss that is a MEAN over the local batch rows (equal shards): returns (loss', logs') whose VALUES are the
    global means (one all-reduce of the stacked scalars) while d loss'/d params = d loss / d params / world -- so that
    the SUM all-reduce of ``FlatParams.allreduce_grads`` yields the single-process gradient.  No-op without
    torch.distributed."""
    rank, ws = world()
    if ws == 1:
        return loss, logs
    vals = torch.stack([loss.detach().float().reshape(())] + [torch.as_tensor(v, dtype=torch.float32, device=loss.device).detach().reshape(()) for v in logs])
    vals = vals / ws
    dist.all_reduce(vals)
    local = loss / ws
    out = local + (vals[0] - local.detach())
    return out, [vals[i + 1] for i in range(len(logs))]


def globalize_sum_loss(loss: torch.Tensor) -> torch.Tensor:
    """For a loss already normalised by the GLOBAL batch (this rank holds a partial sum): value = all-reduced sum,
    gradient = the local one."""
    rank, ws = world()
    if ws == 1:
        return loss
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    return loss + (tot - loss.detach())


class FlatParams:
    """Views of all trainable parameters (and their grads) inside two flat fp32 buffers, so that the
    gradient all-reduce and the fused Adam step are ONE collective / ONE kernel per step."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            k = p.numel()
            self.flat[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + k].view_as(p.data)
            p.grad = self.grad[o:o + k].view_as(p.data)
            o += k
        self.step_count = 0
        # device-resident optimizer state for CUDA-graph replay: (steps taken so far, learning rate)
        self.state = torch.zeros(2, dtype=torch.float32, device=dev)
        self._lr_on_device = None

    def zero_grad(self):
        from . import ops
        with ops._Prof("zero_grad"):
            self.grad.zero_()
        o = 0
        for p in self.params:           # autograd may have replaced .grad; re-point it at the flat buffer
            k = p.numel()
            p.grad = self.grad[o:o + k].view_as(p.data)
            o += k

    def allreduce_grads(self, async_op: bool = False):
        """SUM over ranks (losses are normalised by the global batch, so no division here).  ``async_op=True`` only
        LAUNCHES the collective (NCCL's stream) and returns a handle: call ``wait_grads(handle)`` before the optimizer
        step -- other work (e.g. the probe's forward / backward) can run under the all-reduce."""
        from . import ops
        rank, ws = world()
        if ws > 1:
            with ops._Prof("grad_allreduce"):
                return dist.all_reduce(self.grad, async_op=async_op)
        return None

    @staticmethod
    def wait_grads(handle):
        from . import ops
        if handle is not None:
            with ops._Prof("grad_allreduce_wait"):
                handle.wait()

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, capturable=False):
        """Fused flat-buffer Adam / AdamW.  ``capturable=True`` keeps the step count and the learning rate in device
        memory (``self.state``) so the call can sit inside a captured CUDA graph; set a new learning rate with
        ``set_lr`` between replays."""
        from . import ops
        with ops._Prof("adam"):
            if capturable:
                if self._lr_on_device is None:
                    self.set_lr(lr)
                ops.adam_step_flat_dev(self.flat, self.grad, self.m, self.v, self.state, betas, eps, weight_decay, decoupled)
                return
            self.step_count += 1
            ops.adam_step_flat(self.flat, self.grad, self.m, self.v, lr, self.step_count, betas, eps, weight_decay, decoupled)

    def set_lr(self, lr: float):
        """Write the learning rate (and the host-side step count) into the device state; not capturable."""
        self.state.copy_(torch.tensor([float(self.step_count), float(lr)], dtype=torch.float32))
        self._lr_on_device = float(lr)


def init_process_group_nccl(device, high_priority: bool = True) -> None:
    """``dist.init_process_group("nccl")`` with the collectives on a HIGH-PRIORITY stream.  The InfoNCE kernels queue
    several CTAs per SM (one resident at a time: they take the whole shared memory); at normal priority the CTAs of an
    all-gather issued meanwhile only get an SM once that whole queue has been dispatched, so the gather the next
    critic call waits for sits behind a full kernel.  At high priority the block scheduler hands the next free SM to
    the collective, which then overlaps the tiles.  (Captured into the step's CUDA graph as the node's priority.)"""
    opts = None
    if high_priority:
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:  # noqa: BLE001
            opts = None
    if opts is not None:
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=device)


class GraphedStep:
    """A whole training step (forward, backward, NCCL collectives, fused optimizer) captured ONCE into a CUDA graph
    and replayed: one graph launch per step instead of ~600 kernel launches, which is what keeps an 8-GPU strong-
    scaling run (a few ms of GPU work per launch-bound phase) from being host-bound.  ``fn`` must read its inputs
    from fixed device buffers and must not synchronise; anything that changes per step through HOST scalars (RNG
    seeds, step counts, learning rates) has to live in device memory or run outside the graph."""

    def __init__(self, fn, warmup: int = 2, pool=None, stream=None, on_capture=None):
        # Run the warm-up and the capture on ONE non-default stream.  Everything that touched the parameters
        # before (eager steps, optimizer state) should have run on a non-default stream too: autograd replays
        # each AccumulateGrad on the stream its node was created on, and a node created on the legacy default
        # stream makes capture fail ("legacy stream would depend on a capturing stream").
        side = stream if stream is not None else torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if on_capture is not None:        # e.g. switch the event-record phase markers on for the captured pass only
            on_capture(True)
        try:
            with torch.cuda.graph(self.graph, pool=pool, stream=side):
                self.out = fn()
        finally:
            if on_capture is not None:
                on_capture(False)

    def pool(self):
        return self.graph.pool()

    def __call__(self):
        self.graph.replay()
        return self.out
