"""GPU-box probe: device generator and draw_lanes_batch / offset indicator at config 2's size (256 x 1080p) against the
cv2 calls of the reference on the host.  Prints one JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from draw_util import cv2_draw_lanes, cv2_offset_indicator  # noqa: E402
from multimodal_autonomous_driving_perception_and_planning_b200 import (LaneDetector, OverlayRenderer,  # noqa: E402
                                                                        SyntheticDataGenerator, draw_lanes_batch)

n, w, h = 256, 1920, 1080
out = {"frames": n, "resolution": [w, h]}
gen = SyntheticDataGenerator(w, h)
gen.generate_batch_device(8, start_frame=0)                       # warm-up (library load, static layers)
torch.cuda.synchronize()
t0 = time.perf_counter()
frames, ms = gen.generate_batch_device(n, start_frame=0, return_ms=True)
torch.cuda.synchronize()
out["generator_device_wall_ms"] = (time.perf_counter() - t0) * 1e3
out["generator_device_kernel_ms"] = ms
t0 = time.perf_counter()
host = SyntheticDataGenerator(w, h).generate_batch(16, start_frame=0)
out["generator_cv2_host_ms_per_frame"] = (time.perf_counter() - t0) * 1e3 / 16
assert np.array_equal(host, frames[:16].cpu().numpy())

det = LaneDetector()
lanes = det.detect_batch(frames)
offs = [det.get_lane_center_offset(w, l, r) for l, r in lanes]
ov = OverlayRenderer()
work = frames.clone()
draw_lanes_batch(work, lanes)                                      # warm-up
ov.draw_lane_offset_indicator_batch(work, offs)
best = {"lanes_wall": 1e9, "lanes_dev": 1e9, "ind_wall": 1e9}
for _ in range(3):
    work.copy_(frames)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, ms = draw_lanes_batch(work, lanes, return_ms=True)
    t1 = time.perf_counter()
    ov.draw_lane_offset_indicator_batch(work, offs)
    t2 = time.perf_counter()
    best["lanes_wall"] = min(best["lanes_wall"], (t1 - t0) * 1e3)
    best["lanes_dev"] = min(best["lanes_dev"], ms)
    best["ind_wall"] = min(best["ind_wall"], (t2 - t1) * 1e3)
out["draw_lanes_batch_wall_ms"] = best["lanes_wall"]
out["draw_lanes_batch_device_ms"] = best["lanes_dev"]
out["offset_indicator_batch_wall_ms"] = best["ind_wall"]
hf = frames[:16].cpu().numpy()
t0 = time.perf_counter()
for i in range(16):
    l, r = lanes[i]
    a = cv2_draw_lanes(hf[i].copy(), None if l is None else l.points, None if r is None else r.points)
    cv2_offset_indicator(a, offs[i])
out["cv2_host_draw_ms_per_frame"] = (time.perf_counter() - t0) * 1e3 / 16
out["draw_frames_per_s_device_wall"] = n / ((best["lanes_wall"] + best["ind_wall"]) * 1e-3)
out["draw_frames_per_s_cv2_one_core"] = 1e3 / out["cv2_host_draw_ms_per_frame"]
print(json.dumps(out))
