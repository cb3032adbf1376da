"""N>1 host logic on CPU: view sharding and frame gathering with a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_views_matches_distributed_sampler():
    from pgdvs_b200.dist import shard_views
    from torch.utils.data import DistributedSampler
    for n, world in [(144, 8), (80, 8), (7, 2), (5, 4), (3, 4)]:
        seen = []
        for r in range(world):
            mine = shard_views(n, r, world, pad=True)
            ref = list(DistributedSampler(range(n), num_replicas=world, rank=r, shuffle=False))
            assert mine == ref, (n, world, r)
            seen += shard_views(n, r, world)
        assert sorted(seen) == list(range(n))  # without padding: a partition of the views
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def _worker(rank, world, port, n_views, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pgdvs_b200.dist import gather_frames, shard_views
    mine = shard_views(n_views, rank, world, pad=True)
    # a "frame" that encodes its global view index
    local = torch.stack([torch.full((2, 3, 3), float(v)) for v in mine])
    out = gather_frames(local, n_views, dst=0)
    if rank == 0:
        assert out.shape == (n_views, 2, 3, 3)
        assert out[:, 0, 0, 0].tolist() == [float(v) for v in range(n_views)]
        ok.value = 1
    else:
        assert out is None
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [6, 7])
def test_gather_frames_gloo_world2(n_views):
    ctx = mp.get_context("spawn")
    ok = ctx.Value("i", 0)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ok.value == 1
