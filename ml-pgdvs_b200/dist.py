"""Multi-GPU plumbing: shard target views over ranks, gather rendered frames.

The path is embarrassingly parallel over target views — exactly how the reference
parallelises (one process per GPU, `DistributedSampler(shuffle=False)`, 1 view per GPU per
step: /root/reference/pgdvs/engines/trainer_pgdvs.py:290-306, scripts/benchmark.sh:316).
There is no exchange inside the path, so no data-path collective; NCCL (over NVLink, P2P left
enabled — the reference sets NCCL_P2P_DISABLE=1, scripts/benchmark.sh:36) is used only to
gather finished frames on one rank, e.g. for the video writer
(/root/reference/pgdvs/engines/visualizer_pgdvs.py:160-177 globs per-rank PNGs from disk instead).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int, pad: bool = False) -> List[int]:
    """View indices of `rank`: v -> rank (v mod world_size), like DistributedSampler(shuffle=False).
    With pad=True the tail is padded by wrapping around (DistributedSampler's drop_last=False
    behaviour) so that every rank gets the same count."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    idx = list(range(n_views))
    if pad and n_views % world_size:
        total = ((n_views + world_size - 1) // world_size) * world_size
        idx += idx[: total - n_views]
    return idx[rank::world_size]


def gather_frames(local: torch.Tensor, n_views: int, dst: int = 0, group=None,
                  out_list: Optional[List[torch.Tensor]] = None) -> Optional[torch.Tensor]:
    """Gather per-rank frames [V_local, ...] (sharded with shard_views(..., pad=True)) on `dst`
    and restore the global view order.  Returns [n_views, ...] on dst, None elsewhere.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors; used by the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local[:n_views]
    if rank == dst:
        bufs = out_list if out_list is not None else [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, bufs, dst=dst, group=group)
        # rank r holds views r, r + world, ...  ->  interleave
        stacked = torch.stack(bufs, dim=1)  # [V_local, world, ...]
        return stacked.reshape((-1,) + tuple(local.shape[1:]))[:n_views]
    dist.gather(local, None, dst=dst, group=group)
    return None


# ------------------------------------------------------------------ copy-engine frame sink
class _DevMem:
    """A raw device allocation viewed as a torch tensor (through __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerFrameSink:
    """Gather of finished frames on rank `dst` WITHOUT kernels: `dst` exports a device buffer over
    CUDA IPC (libpgdvs_b200: pgdvs_ipc_*), every rank writes its frames into its slot with a
    stream-ordered copy-engine copy over NVLink (pgdvs_copy_async).  Slots are double-buffered by
    step parity; `commit()` is a one-element all-reduce on the copy stream that orders the writes of
    step s before anybody's step s + 2 and tells `dst` that step s has landed.

        sink = PeerFrameSink((V, H, W, 3), torch.uint8, device)       # collective: all ranks
        sink.push(frames_u8, step)                                     # every rank, on sink.stream
        sink.commit()
        sink.frames(step)      # dst only, valid once sink.stream has passed commit(): [world, V, H, W, 3]

    One process per GPU, one node (cudaIpc handles do not cross nodes)."""

    def __init__(self, shape, dtype, device, dst: int = 0, group=None):
        import ctypes

        import numpy as np

        from . import _cabi
        self._cabi, self._ct = _cabi, ctypes
        self.group, self.dst = group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        self.shape, self.dtype = tuple(shape), dtype
        self.slot_bytes = int(np.prod(self.shape)) * torch.empty((), dtype=dtype).element_size()
        self.slot_bytes = (self.slot_bytes + 255) & ~255
        total = 2 * self.world * self.slot_bytes
        self.stream = torch.cuda.Stream(device=self.device)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        L = _cabi.lib()
        handle = (ctypes.c_ubyte * _cabi.IPC_HANDLE_BYTES)()
        self._owned = self._mapped = None
        with torch.cuda.device(self.device):
            if self.rank == dst:
                p = ctypes.c_void_p()
                _cabi.check(L.pgdvs_ipc_alloc(total, ctypes.byref(p), handle), "pgdvs_ipc_alloc")
                self._owned = p.value
            box = [bytes(handle) if self.rank == dst else None]
            dist.broadcast_object_list(box, src=dst, group=group)
            if self.rank == dst:
                self.base = self._owned
            else:
                h = (ctypes.c_ubyte * _cabi.IPC_HANDLE_BYTES).from_buffer_copy(box[0])
                p = ctypes.c_void_p()
                _cabi.check(L.pgdvs_ipc_open(h, ctypes.byref(p)), "pgdvs_ipc_open")
                self._mapped = self.base = p.value
        self._view = None
        if self.rank == dst:
            self._view = torch.as_tensor(_DevMem(self.base, total), device=self.device)
        dist.barrier(group=group)

    def _slot(self, step: int, rank: int) -> int:
        return self.base + ((step & 1) * self.world + rank) * self.slot_bytes

    def push(self, frames: torch.Tensor, step: int, after: Optional[torch.cuda.Event] = None):
        """Copy this rank's frames of `step` into its slot on `dst` (on self.stream, after `after`)."""
        if tuple(frames.shape) != self.shape or frames.dtype != self.dtype or not frames.is_contiguous():
            raise ValueError(f"frames must be contiguous {self.dtype} {self.shape}")
        if after is not None:
            self.stream.wait_event(after)
        frames.record_stream(self.stream)
        with torch.cuda.device(self.device):
            self._cabi.check(self._cabi.lib().pgdvs_copy_async(
                self._slot(step, self.rank), frames.data_ptr(), frames.numel() * frames.element_size(),
                self.stream.cuda_stream), "pgdvs_copy_async")

    def commit(self):
        """Order this step's writes before the slot is reused (one float all-reduce on self.stream)."""
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self._flag, group=self.group)

    def frames(self, step: int) -> torch.Tensor:
        """dst only: [world, *shape] view of the frames of `step` (rank-major; view v of the whole job is
        frames[v % world, v // world] under shard_views)."""
        if self.rank != self.dst:
            raise RuntimeError("frames() is only available on the destination rank")
        n = self.slot_bytes
        off = (step & 1) * self.world * n
        per = int(torch.tensor(self.shape).prod()) * torch.empty((), dtype=self.dtype).element_size()
        v = self._view[off:off + self.world * n].view(self.world, n)[:, :per]
        return v.view(self.dtype).view((self.world,) + self.shape)

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        L = self._cabi.lib()
        with torch.cuda.device(self.device):
            if self._mapped is not None:
                L.pgdvs_ipc_close(self._mapped)
                self._mapped = None
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            if self._owned is not None:
                self._view = None
                L.pgdvs_ipc_free(self._owned)
                self._owned = None


def bind_rank_to_cores(local_rank: int, local_world: int, device_index: Optional[int] = None) -> dict:
    """Pin this process to its share of the host cores: the cores local to its GPU's NUMA node when
    sysfs tells (/sys/bus/pci/devices/<bdf>/local_cpulist), split evenly among the ranks whose GPUs
    report the same list; otherwise an even split of all cores.  Pinned host buffers allocated
    afterwards are first-touched on that node.  Returns what was done (for the bench line)."""
    import os
    info = {"bound": False}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        dev = device_index if device_index is not None else local_rank
        lists = []
        for d in range(local_world):
            bdf = None
            try:
                pr = torch.cuda.get_device_properties(d)
                bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
                    lists.append(f.read().strip())
            except Exception:
                lists.append(None)
        mine = lists[dev] if dev < len(lists) else None
        if mine:
            local = []
            for part in mine.split(","):
                a, _, b = part.partition("-")
                local += list(range(int(a), int(b or a) + 1))
            local = [c for c in local if c in allowed]
            if local:
                cpus = local
                info["numa_cpulist"] = mine
        sharers = [d for d in range(local_world) if lists[d] == (lists[dev] if dev < len(lists) else None)]
        k, n = sharers.index(dev) if dev in sharers else local_rank, max(len(sharers), 1)
        per = max(len(cpus) // n, 1)
        share = cpus[k * per:(k + 1) * per] or cpus
        os.sched_setaffinity(0, share)
        torch.set_num_threads(max(1, min(len(share), 8)))
        info.update(bound=True, cores=[share[0], share[-1]], n_cores=len(share))
    except Exception as e:  # not fatal: the ranks just stay unbound
        info["error"] = repr(e)
    return info
