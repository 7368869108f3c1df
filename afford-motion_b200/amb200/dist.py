"""Multi-GPU plumbing for batch-sharded sampling: one process per GPU, `torch.distributed` (NCCL on GPUs,
gloo in CPU tests) used ONLY for rendezvous, barriers and timing reductions — sampling needs no data-path collective
because samples are independent in eval mode.  Results are rank-count invariant because every rank addresses its Philox
streams by GLOBAL sample index (`diffusion.sample_offset`)."""
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: Optional[str] = None, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of the global batch owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    """Timing reduction: the job is as slow as its slowest rank."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_samples(local: torch.Tensor, global_batch: int) -> Optional[List[torch.Tensor]]:
    """Collect every rank's samples on rank 0 (host side, after sampling; not on the timed path)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard(global_batch, r, world) for r in range(world)]
    if len({e - s for s, e in sizes}) == 1:
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(bufs, local)
    else:  # ragged shards: object gather (host side, small)
        bufs = [None] * world
        dist.all_gather_object(bufs, local.cpu())
    return bufs if rank == 0 else None


def allreduce_flat_(bufs: List[torch.Tensor], group=None) -> int:
    """In-place SUM all-reduce of flat gradient buffers (training exchange step); returns the world size the
    caller divides by (amb200.optim.FusedAdamW folds 1/world into its update kernel).  No-op for a single process."""
    if not dist.is_initialized():
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        for b in bufs:
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    return world


class ChunkedGradExchange:
    """Gradient all-reduce overlapped with backward (the reference gets the overlap from DistributedDataParallel's
    buckets, train_ddp.py:63-65).  Every flat gradient buffer is cut into `chunks` contiguous pieces of whole parameters; a
    post-accumulate-grad hook on each parameter counts its piece down, and the piece's SUM all-reduce is launched
    asynchronously the moment its last gradient has been accumulated (NCCL runs it on its own stream, under the rest of
    backward).  Backward produces gradients in roughly reverse registration order, so the LAST piece goes first.
    `begin()` arms the hooks for the next backward — only the FINAL backward of a step: an all-reduce is linear, so a piece
    that was reduced early must not receive further local gradients.  `finish()` launches the pieces that never completed
    (parameters that received no gradient: `find_unused_parameters=True` in the reference), waits, and returns the world size.
    `flats`: [(flat_grad, params, offsets, numel)] with every p.grad a view of flat_grad at its offset."""

    def __init__(self, flats, chunks: int = 3, group=None):
        self.group, self.armed, self.chunks, self.handles = group, False, [], []
        for g, ps, offs, n in flats:
            per = (n + chunks - 1) // chunks
            lo_i = 0
            while lo_i < len(ps):
                hi_i = lo_i
                while hi_i < len(ps) and offs[hi_i] - offs[lo_i] < per:
                    hi_i += 1
                hi = offs[hi_i] if hi_i < len(ps) else n
                self.chunks.append(dict(view=g[offs[lo_i]:hi], params=ps[lo_i:hi_i], pending=hi_i - lo_i, work=None, early=False))
                lo_i = hi_i
        for c in self.chunks:
            for prm in c["params"]:
                self.handles.append(prm.register_post_accumulate_grad_hook(self._hook(c)))

    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _hook(self, c):
        def hook(_param):
            if not self.armed:
                return
            c["pending"] -= 1
            if c["pending"] == 0 and self._world() > 1:
                c["work"] = dist.all_reduce(c["view"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                c["early"] = True
        return hook

    def begin(self):
        for c in self.chunks:
            c["work"], c["pending"], c["early"] = None, len(c["params"]), False
        self.armed = True

    def finish(self) -> int:
        self.armed = False
        world = self._world()
        if world > 1:
            for c in self.chunks:
                if c["work"] is None:
                    c["work"] = dist.all_reduce(c["view"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            for c in self.chunks:
                c["work"].wait()   # the current stream waits for the collective's stream
                c["work"] = None
        return world

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []
