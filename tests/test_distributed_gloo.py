"""N > 1 path on CPU: world_size-2 gloo processes shard the camera streams and gather the packed
per-frame records on rank 0 (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_autonomous_driving_perception_and_planning_b200 import _native
from multimodal_autonomous_driving_perception_and_planning_b200.distributed import (
    RecordGatherer, gather_records, merge_stream_major, streams_of_rank)


def test_stream_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        owned = [streams_of_rank(8, world, r) for r in range(world)]
        assert sorted(s for o in owned for s in o) == list(range(8))
        assert all(len(o) == 8 // world for o in owned)
    assert streams_of_rank(3, 2, 1) == [1]
    with pytest.raises(ValueError):
        streams_of_rank(8, 2, 2)


def _fake_records(stream, t):
    r = np.zeros(t, _native.RECORD_DTYPE)
    r["n_segments"] = stream * 1000 + np.arange(t)
    r["offset"] = stream + np.arange(t) / 100.0
    r["side"]["coeffs"][:, 0, 0] = stream
    return r


def _worker(rank, world, port, n_streams, t, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = streams_of_rank(n_streams, world, rank)
    local = np.concatenate([_fake_records(s, t) for s in mine]) if mine else np.zeros(0, _native.RECORD_DTYPE)
    got = gather_records(local, dst=0)
    if rank == 0:
        merged = merge_stream_major(got, n_streams, t, world)
        q.put((merged["n_segments"].tolist(), merged["offset"].tolist(), [len(g) for g in got]))
    else:
        assert got is None
    if n_streams % world == 0:                      # fixed-size steady-state gather gives the same bytes
        fixed = RecordGatherer(len(local), torch.device("cpu")).gather(local)
        if rank == 0:
            assert all(np.array_equal(a, b) for a, b in zip(fixed, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams,t", [(8, 5), (3, 4)])
def test_gather_records_world_size_2(n_streams, t):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, t, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    segs, offs, lens = q.get()
    assert lens == [len(streams_of_rank(n_streams, 2, r)) * t for r in range(2)]
    for s in range(n_streams):
        assert segs[s] == [s * 1000 + i for i in range(t)]
        assert np.allclose(offs[s], s + np.arange(t) / 100.0)
