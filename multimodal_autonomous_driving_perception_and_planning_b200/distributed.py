"""Multi-GPU plumbing for the lane-detection path: one process per GPU, streams sharded, results gathered.

The path shards with no exchange: every stage before the temporal smoothing is
per frame, and the smoothing state is per camera stream, so whole streams are assigned to ranks and
nothing crosses GPUs while frames are processed.  The only collective is one gather of the packed
per-frame ``lane_record`` array to rank 0 (NCCL over NVLink on the GPU box; any torch.distributed
backend works, the CPU tests use gloo).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from ._native import RECORD_DTYPE


def streams_of_rank(n_streams: int, world_size: int, rank: int) -> List[int]:
    """Camera streams owned by ``rank``: stream s -> rank s % world_size (BASELINE config 3:
    8 cameras over 1/2/4/8 GPUs gives 8/4/2/1 whole streams per GPU)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    return [s for s in range(n_streams) if s % world_size == rank]


def gather_records(local: np.ndarray, dst: int = 0, device=None, group=None) -> Optional[List[np.ndarray]]:
    """Gather every rank's ``RECORD_DTYPE`` array on ``dst`` (ragged lengths allowed).

    Returns the list indexed by rank on ``dst`` and ``None`` elsewhere.  Two collectives: the
    lengths (all_gather of one int64) and the padded byte payload (gather)."""
    import torch
    import torch.distributed as dist
    if local.dtype != RECORD_DTYPE:
        raise TypeError("gather_records expects a RECORD_DTYPE array")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    item = RECORD_DTYPE.itemsize
    cap = max(max(counts), 1) * item
    payload = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if local.shape[0]:
        raw = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1))
        payload[: raw.numel()] = raw.to(dev)
    bufs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [b[: counts[r] * item].cpu().numpy().view(RECORD_DTYPE).copy() for r, b in enumerate(bufs)]


def merge_stream_major(per_rank: Sequence[np.ndarray], n_streams: int, frames_per_stream: int,
                       world_size: int) -> np.ndarray:
    """Reassemble rank-gathered records into [n_streams, frames_per_stream] order, given that each rank
    processed its streams (``streams_of_rank``) back to back, every stream in temporal order."""
    out = np.zeros((n_streams, frames_per_stream), RECORD_DTYPE)
    for rank, recs in enumerate(per_rank):
        mine = streams_of_rank(n_streams, world_size, rank)
        if recs.shape[0] != len(mine) * frames_per_stream:
            raise ValueError(f"rank {rank} returned {recs.shape[0]} records for {len(mine)} streams")
        for i, s in enumerate(mine):
            out[s] = recs[i * frames_per_stream:(i + 1) * frames_per_stream]
    return out


class _DeviceBytes:
    """Zero-copy view of raw device memory for ``torch.as_tensor`` (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class RecordGatherer:
    """Fixed-size gather for steady-state serving: every rank contributes exactly ``n`` records per
    call, buffers are allocated once, one collective per call.

    ``gather(local)`` takes the host records (staged through one of two pinned buffers; the copy of call k
    is fenced by an event before buffer k % 2 is rewritten by call k + 2).  ``gather_device(ptr)`` takes the
    device copy the native context keeps (``LaneContext.records_device_ptr()``): the records go from the
    context's buffer straight into the collective, with no host round trip.  On the destination rank both return
    the per-rank device buffers (``to_host=False``) or decoded ``RECORD_DTYPE`` arrays."""

    def __init__(self, n: int, device, dst: int = 0, group=None):
        import torch
        import torch.distributed as dist
        self.n, self.dst, self.group, self.device = n, dst, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = nbytes = n * RECORD_DTYPE.itemsize
        cuda = device.type == "cuda"
        self.stage = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() if cuda else torch.empty(nbytes, dtype=torch.uint8)
                      for _ in range(2)]
        self.payload = [torch.empty(nbytes, dtype=torch.uint8, device=device) for _ in range(2)]
        self.copied = [None, None]                   # CUDA event after the H2D copy out of stage[i]
        self.calls = 0
        self.bufs = [torch.empty(nbytes, dtype=torch.uint8, device=device) for _ in range(self.world)] \
            if self.rank == dst else None

    def _finish(self, payload, to_host):
        import torch.distributed as dist
        dist.gather(payload, self.bufs, dst=self.dst, group=self.group)
        if self.rank != self.dst:
            return None
        if not to_host:
            return self.bufs
        return [b.cpu().numpy().view(RECORD_DTYPE).copy() for b in self.bufs]

    def gather(self, local: np.ndarray, to_host: bool = True):
        if local.shape[0] != self.n or local.dtype != RECORD_DTYPE:
            raise ValueError("RecordGatherer.gather expects exactly n RECORD_DTYPE records")
        k = self.calls & 1
        self.calls += 1
        if self.copied[k] is not None:
            self.copied[k].synchronize()             # the DMA that last read this staging buffer has finished
        self.stage[k].numpy()[:] = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
        self.payload[k].copy_(self.stage[k], non_blocking=True)
        if self.device.type == "cuda":
            import torch
            self.copied[k] = torch.cuda.Event()
            self.copied[k].record(torch.cuda.current_stream(self.device))
        return self._finish(self.payload[k], to_host)

    def gather_device(self, records_ptr: int, to_host: bool = False):
        """``records_ptr``: device address of ``n`` packed ``lane_record``s on ``self.device`` (the batch has been collected,
        so its records are complete).  The collective is enqueued on torch's current stream; call
        ``LaneContext.fence_records(torch.cuda.current_stream().cuda_stream)`` right after, so that the kernel that next
        writes that result slot -- and only that kernel -- waits for it."""
        import torch
        src = torch.as_tensor(_DeviceBytes(records_ptr, self.nbytes), device=self.device)
        return self._finish(src, to_host)
