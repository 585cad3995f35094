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
    """Gather a row-sharded tensor [n_r, ...] from every rank into [sum n_r, ...] on every rank, in rank order, with ONE
    `all_gather_into_tensor` (a single NCCL all-gather into one contiguous buffer; no per-rank tensor list, no cat).
    With `total` given the shards are the contiguous `shard_bounds(total, rank, world)` split -- every rank but the trailing
    ones holds ceil(total/world) rows -- so the padded concatenation IS the result and no size exchange is needed.
    Without `total` the shard sizes are exchanged first (ragged shards in any distribution)."""
    rank, world = world_info()
    if world == 1:
        return t
    tail = tuple(t.shape[1:])
    if total is not None:
        per = (total + world - 1) // world
        lo, hi = shard_bounds(total, rank, world)
        assert t.shape[0] == hi - lo, "all_gather_rows(total=%d): rank %d holds %d rows, expected %d" % (total, rank, t.shape[0], hi - lo)
        if per == 0:
            return t.new_empty((0,) + tail)
        send = t if t.shape[0] == per else torch.cat([t, t.new_zeros((per - t.shape[0],) + tail)], dim=0)
        out = t.new_empty((world * per,) + tail)
        dist.all_gather_into_tensor(out, send.contiguous())
        return out[:total]
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes_t = torch.empty(world, device=t.device, dtype=torch.int64)
    dist.all_gather_into_tensor(sizes_t, n)
    sizes = [int(x) for x in sizes_t.tolist()]
    m = max(sizes)
    if m == 0:
        return t.new_empty((0,) + tail)
    send = t if t.shape[0] == m else torch.cat([t, t.new_zeros((m - t.shape[0],) + tail)], dim=0)
    buf = t.new_empty((world * m,) + tail)
    dist.all_gather_into_tensor(buf, send.contiguous())
    if all(x == m for x in sizes):
        return buf
    return torch.cat([buf[r * m:r * m + x] for r, x in enumerate(sizes)], dim=0)


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
