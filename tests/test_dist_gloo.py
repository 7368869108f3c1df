"""world_size-2 gloo tests of the multi-GPU host plumbing (shard bookkeeping, timing max-reduce, result gather)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "afford-motion_b200"))
    from amb200 import dist as D
    r, w = D.init("gloo")
    s, e = D.shard(33, r, w)
    slow = D.max_over_ranks(10.0 + 5.0 * r)
    D.barrier()
    local = torch.arange(s, e, dtype=torch.float32).view(-1, 1).repeat(1, 3)
    if (e - s) != 33 // w:  # uneven shards: equalise for the tensor all_gather used on the happy path
        local = local[: 33 // w]
    got = D.gather_samples(local, (33 // w) * w)
    # training exchange step: SUM all-reduce of a flat gradient buffer + the 1/world scale the fused AdamW folds in
    flat = [torch.full((10,), float(r + 1)), torch.arange(4, dtype=torch.float32) * (r + 1)]
    wsz = D.allreduce_flat_(flat)
    mean_ok = wsz == w and torch.allclose(flat[0] / wsz, torch.full((10,), 1.5)) and torch.allclose(flat[1] / wsz, torch.arange(4.0) * 1.5)
    q.put((r, (s, e), slow, None if got is None else torch.cat(got).tolist(), bool(mean_ok)))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, sh0, slow0, g0, ok0), (r1, sh1, slow1, g1, ok1) = res
    assert ok0 and ok1                                     # flat gradient all-reduce gives the mean on every rank
    assert sh0 == (0, 17) and sh1 == (17, 33)            # contiguous, exhaustive, sizes differ by <= 1
    assert slow0 == slow1 == 15.0                          # max over ranks on every rank
    assert g1 is None and len(g0) == 32 and g0[0] == [0.0, 0.0, 0.0] and g0[16][0] == 17.0


def test_shard_partitions_any_batch():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "afford-motion_b200"))
    from amb200.dist import shard
    for gb in (1, 7, 16, 64, 256):
        for world in (1, 2, 4, 8):
            spans = [shard(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _overlap_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "afford-motion_b200"))
    import torch.distributed as dist
    from amb200 import dist as D
    D.init("gloo")
    torch.manual_seed(0)
    # five parameters whose gradients are views of ONE flat buffer, as amb200.optim.FusedAdamW lays them out; `unused` gets no gradient
    shapes = [(7, 3), (5,), (4, 4), (9,), (2, 6)]
    offs, n = [], 0
    for sh in shapes:
        offs.append(n)
        n += (int(torch.tensor(sh).prod()) + 3) // 4 * 4
    flat = torch.zeros(n)
    ps = []
    for sh, o in zip(shapes, offs):
        prm = torch.nn.Parameter(torch.randn(*sh))
        prm.grad = flat[o:o + prm.numel()].view(sh)
        ps.append(prm)
    ex = D.ChunkedGradExchange([(flat, ps, offs, n)], chunks=3)
    x = torch.full((3,), float(rank + 1))
    def loss_fn():
        return (ps[0] @ x).sum() * (rank + 1) + (ps[1] ** 2).sum() + (ps[2] * (rank + 2)).sum() + (ps[4].sum() * 3.0)   # ps[3] unused
    # reference: plain backward + one flat all-reduce
    loss_fn().backward()
    ref = flat.clone()
    dist.all_reduce(ref)
    flat.zero_()
    # overlapped: hooks launch complete pieces during backward, finish() reduces the piece holding the unused parameter
    ex.begin()
    loss_fn().backward()
    early = [c["early"] for c in ex.chunks]
    wsz = ex.finish()
    ok = wsz == world and torch.allclose(flat, ref) and any(early) and not all(early)
    # a second step re-arms cleanly; an un-armed backward launches nothing
    flat.zero_()
    loss_fn().backward()
    quiet = all(c["work"] is None for c in ex.chunks)
    q.put((rank, bool(ok), bool(quiet), len(ex.chunks)))
    dist.destroy_process_group()


def test_chunked_grad_exchange_overlaps_and_matches_flat_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for _, ok, quiet, nchunks in res:
        assert ok and quiet and nchunks == 3
