"""Host-side multi-process logic on CPU: world_size-2 gloo run of the tile packing + all-gather
used by the N>1 path (the render itself is GPU-only and is covered by the gpu tests)."""
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


def _worker(rank, world, port, F, R):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__

    __graft_entry__.build()
    from neural_graph_mapping_b200 import distributed as D

    assert D.world_info() == (world, rank)
    n = F * R
    local = torch.empty(D.FLOATS_PER_RAY * n)
    rgbd, cvar, dvar, term = D.packed_views(local, n)
    rgbd.fill_(rank + 0.25)
    cvar.fill_(rank + 0.5)
    dvar.fill_(rank + 0.75)
    term.copy_(torch.arange(n, dtype=torch.float32) + 1000 * rank)
    buf = D.gather_tiles(local)
    assert buf.shape == (world, D.FLOATS_PER_RAY * n)
    for r in range(world):
        a, b, c, d = D.packed_views(buf[r], n)
        assert torch.all(a == r + 0.25) and torch.all(b == r + 0.5) and torch.all(c == r + 0.75)
        assert torch.equal(d, torch.arange(n, dtype=torch.float32) + 1000 * r)
    # the asynchronous form: this rank's tile at once, the full buffer after wait()
    pend = D.gather_tiles(local, async_op=True)
    assert torch.equal(pend.local, local)
    assert torch.equal(pend.wait(), buf) and torch.equal(pend.wait(), buf)  # wait() is idempotent
    # one jitter seed for all ranks (rank 0's draw)
    torch.manual_seed(100 + rank)
    seed = D.shared_seed("cpu")
    seeds = [None] * world
    dist.all_gather_object(seeds, seed)
    assert len(set(seeds)) == 1
    # balanced contiguous field shards cover [0, F) exactly once
    cover = []
    for r in range(world):
        s, e = D.shard_range(F, world, r)
        cover += list(range(s, e))
    assert cover == list(range(F))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gather_tiles_gloo_world2():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 3, 5), nprocs=2, join=True)


def test_shard_range_uneven():
    from neural_graph_mapping_b200 import distributed as D

    assert [D.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert D.shard_range(256, 8, 7) == (224, 256)
    assert D.world_info() == (1, 0)


def test_leading_slice_only_touches_whole_batch_inputs():
    """render_rays_sharded / render_rays_split slice per-ray inputs given for the WHOLE batch and pass everything
    else through (scalars, a shared (4, 4) pose, inputs the caller already sliced)."""
    from neural_graph_mapping_b200 import distributed as D

    F, R = 6, 10
    near = torch.arange(F * R, dtype=torch.float32).view(F, R)
    jit = torch.rand(F, R, 8)
    c2ws = torch.eye(4).repeat(F, R, 1, 1)
    sl = slice(2, 4)
    assert torch.equal(D._leading_slice(near, (F, R), sl), near[2:4])
    assert torch.equal(D._leading_slice(jit, (F, R), sl), jit[2:4])
    assert D._leading_slice(c2ws, (F, R), sl).shape == (2, R, 4, 4)
    assert D._leading_slice(None, (F, R), sl) is None and D._leading_slice(1.5, (F, R), sl) == 1.5
    pre = jit[2:4]
    assert D._leading_slice(pre, (F, R), sl) is pre            # already a shard: untouched
    rays = (slice(None), slice(5, 10))
    assert torch.equal(D._leading_slice(near, (F, R), rays), near[:, 5:10])


def test_owned_fields_partition_targets_exactly_once():
    """Every target field of a training iteration is updated by exactly one rank (no gradient collective)."""
    from neural_graph_mapping_b200 import distributed as D

    g = torch.Generator().manual_seed(0)
    for num_fields, world in ((256, 8), (37, 4), (5, 8)):
        ids = torch.randperm(num_fields, generator=g)[: min(32, num_fields)]
        owners = torch.stack([D.owned_fields(ids, num_fields, world, r) for r in range(world)]).long()
        assert torch.equal(owners.sum(0), torch.ones(len(ids), dtype=torch.long))
        for r in range(world):
            f0, f1 = D.shard_range(num_fields, world, r)
            assert all(f0 <= int(i) < f1 for i in ids[owners[r].bool()])
