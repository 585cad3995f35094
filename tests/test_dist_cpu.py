"""Host-side sharding / gather logic of the N>1 path on CPU: world_size 2 (and 3, ragged) over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from whmr_b200.dist import all_gather_rows, all_reduce_sum, max_over_ranks, shard_bounds
    lo, hi = shard_bounds(total, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float32)
    per_frame = torch.stack([idx * 2.0, idx + 0.5, idx * idx], dim=1)      # stand-in for [n_r, 3] errors
    g = all_gather_rows(per_frame, total)
    g2 = all_gather_rows(per_frame)            # ragged path: sizes exchanged first
    assert torch.equal(g, g2)
    rag = all_gather_rows(per_frame[: (rank + 1) % 2 * per_frame.shape[0]])   # arbitrary ragged shards (some empty)
    assert rag.shape[0] == sum(((r + 1) % 2) * (shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0]) for r in range(world))
    s = all_reduce_sum(per_frame.sum(0).clone())
    mx = max_over_ranks(float(rank + 1), torch.device("cpu"))
    if rank == 0:
        q.put((g, s, mx))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 35515), (2, 1), (3, 7), (2, 0)])
def test_shard_and_gather(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    g, s, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = torch.arange(total, dtype=torch.float32)
    ref = torch.stack([idx * 2.0, idx + 0.5, idx * idx], dim=1)
    assert torch.equal(g, ref)
    assert torch.allclose(s, ref.sum(0), rtol=1e-6)
    assert mx == float(world)


def test_shard_bounds_cover_exactly_once():
    from whmr_b200.dist import shard_bounds
    for total in (0, 1, 7, 256, 1000, 35515, 65536):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
