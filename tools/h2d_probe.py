#!/usr/bin/env python
"""Host->device ceiling of the GPU box: N processes, one per GPU, each copying the bench's per-step payload
(256 x 1080p BGR frames = 1.59 GB, pinned) to its GPU with plain cudaMemcpyAsync, nothing else running.

    python tools/h2d_probe.py                      # one process, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/h2d_probe.py                         # N processes started together

Prints one JSON line (rank 0): per-GPU GB/s (min / mean / max over ranks) and the aggregate.  This is the roof the
e2e leg of bench.py is measured against (`e2e.pcie_frac`): e2e moves the same bytes through LaneDetector.detect_batch.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=256 * 1080 * 1920 * 3)
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--chunks", type=int, default=4, help="copies per payload (bench.py's detect path uses 4 chunks)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")
    host = torch.empty(args.bytes, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    dst = torch.empty(args.bytes, dtype=torch.uint8, device=dev)
    step = (args.bytes + args.chunks - 1) // args.chunks

    def copy_once():
        for a in range(0, args.bytes, step):
            dst[a:a + step].copy_(host[a:a + step], non_blocking=True)

    for _ in range(3):
        copy_once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        copy_once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = args.bytes * args.reps / dt / 1e9
    vals = [gbs]
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, gbs)
        vals = out
    if rank == 0:
        print(json.dumps({"probe": "pinned H2D, one process per GPU, concurrent", "n_gpus": world,
                          "bytes_per_copy": args.bytes, "chunks": args.chunks, "reps": args.reps,
                          "gbs_per_gpu_min": min(vals), "gbs_per_gpu_mean": sum(vals) / len(vals),
                          "gbs_per_gpu_max": max(vals), "gbs_aggregate": sum(vals),
                          "frames_per_s_equiv_1080p_bgr": sum(vals) * 1e9 / (1080 * 1920 * 3)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
