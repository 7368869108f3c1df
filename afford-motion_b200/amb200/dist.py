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
