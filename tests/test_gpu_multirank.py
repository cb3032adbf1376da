"""Multi-rank GPU test: the union of the per-rank shards equals the single-GPU
render bit for bit.  Two processes shard the views exactly like the reference's
DistributedSampler(shuffle=False) (dist.shard_views), render their shard and gather the frames on
rank 0 (dist.gather_frames).  With two or more GPUs the ranks use one GPU each and NCCL; on a
one-GPU box both ranks render on cuda:0 and the frames travel as CPU tensors over gloo — the
property under test (sharding + order restoration + batch-independence of the kernels) is the same."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _render(wl, views, dev):
    import pgdvs_b200
    pairs, cams = wl.jobs(views)
    out = pgdvs_b200.render_views(pairs, cams, wl.H, wl.W, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                                  static_rgb=wl.static_rgb[list(views)], return_fragments=True, return_u8=True)
    return out


def _worker(rank, world, port, n_gpus, n_views, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    nccl = n_gpus >= world
    dev = torch.device("cuda", rank if nccl else 0)
    torch.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dist import gather_frames, shard_views
    wl = synthetic.make_workload("tiny_track", dev, n_views=n_views, seed=5)  # the same job on every rank
    mine = shard_views(n_views, rank, world, pad=True)
    out = _render(wl, mine, dev)
    got = {}

    def local_idx(o):
        # idx is an index into the packed cloud of the LAUNCH (pytorch3d semantics): make it relative
        # to its own view's first point so that shards and the full batch can be compared
        first = o["first_idx"].to(torch.int32)[:, None, None, None]
        return torch.where(o["idx"] >= 0, o["idx"] - first, o["idx"])

    out["idx"] = local_idx(out)
    for k in ("image", "image_u8", "mask", "idx", "zbuf", "dists"):
        t = out[k].contiguous()
        g = gather_frames(t if nccl else t.cpu(), n_views, dst=0)
        if rank == 0:
            got[k] = g.cpu()
    failure = None
    if rank == 0:
        try:
            full = _render(wl, range(n_views), dev)
            full["idx"] = local_idx(full)
            for k, g in got.items():
                assert g.shape[0] == n_views
                assert torch.equal(g, full[k].cpu()), f"{k}: union of shards differs from the single-GPU render"
            ok.value = 1
        except Exception as e:  # noqa: BLE001  (the other rank must not be left waiting at the barrier)
            failure = e
    dist.barrier()
    dist.destroy_process_group()
    if failure is not None:
        raise failure


@pytest.mark.parametrize("n_views", [6, 7])
def test_union_of_shards_equals_single_gpu_render(n_views):
    n_gpus = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    ok = ctx.Value("i", 0)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_gpus, n_views, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    for p in procs:
        if p.exitcode is None:
            p.terminate()
    assert [p.exitcode for p in procs] == [0, 0]
    assert ok.value == 1
