"""GPU-box soak of the two-stream streaming path: six different 1080p batches (one with dense noise frames) cycled two deep
for many rounds; every collected batch's records must equal the batch-after-batch result byte for byte.  Prints a summary."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, SyntheticDataGenerator

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n, w, h = 128, 1920, 1080
gen = SyntheticDataGenerator(w, h)
batches = []
for b in range(6):
    fr = gen.generate_batch_device(n, start_frame=53 * b)
    if b == 4:
        g = torch.Generator(device="cuda").manual_seed(3)
        fr[:8] = torch.randint(0, 256, fr[:8].shape, dtype=torch.uint8, device="cuda", generator=g)
    batches.append(fr)
det = LaneDetector(max_batch=n, max_segments=4096)
ctx = det._context(h, w, n)
pf, pv = np.zeros((1, 2, 3)), np.zeros((1, 2), np.uint8)
# reference chain: the same cyclic sequence, one batch at a time; the EMA state makes every position of the cycle depend on
# everything before it, so the reference is computed for the whole sequence
total = rounds * len(batches)
want = []
for k in range(total):
    want.append(ctx.detect(batches[k % 6].data_ptr(), n, True, None, 1, pf, pv, 0.7, 1 - 0.7).tobytes())
pf[:] = 0
pv[:] = 0
bad = 0
t0 = time.perf_counter()
ctx.enqueue(batches[0].data_ptr(), n, None, 1, pf, pv, 0.7, 1 - 0.7)
for k in range(total):
    if k + 1 < total:
        ctx.enqueue(batches[(k + 1) % 6].data_ptr(), n, None, 1, None, None, 0.7, 1 - 0.7)
    bad += ctx.collect(pf, pv).tobytes() != want[k]
dt = time.perf_counter() - t0
print(f"soak done: {total} batches of {n} x 1080p two deep, {bad} differ from the batch-after-batch records; {total * n / dt:.0f} frames/s")
det.close()
