"""Multi-GPU plumbing for the hot path: bodies are independent, so the batch is sharded contiguously
by rank and NOTHING is exchanged during compute.  The only collective is the gather of
per-frame metrics / small outputs at the end of a pass (the reference's one explicit collective is a
scalar all_reduce, core/trainer.py:630).  Works with the `nccl` backend on GPUs and with `gloo` on
CPU (tests/test_dist_cpu.py)."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """rank r of R gets rows [r*ceil(B/R), min(B, (r+1)*ceil(B/R)))"""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(t, total=None):
    """Gather a row-sharded tensor [n_r, ...] from every rank into [sum n_r, ...] on every rank, in
    rank order.  Shards may be ragged (the last ranks can be short or empty)."""
    rank, world = world_info()
    if world == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    if total is not None:
        assert out.shape[0] == total, (out.shape, total)
    return out


def all_reduce_sum(t):
    _, world = world_info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(x, device):
    t = torch.tensor([float(x)], device=device)
    _, world = world_info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
