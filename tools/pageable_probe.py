"""e2e of LaneDetector.detect_batch for ordinary (pageable) numpy frames vs pinned ones, 256 x 1080p."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, multi_camera_batch
base = multi_camera_batch(1, 64, 1920, 1080)[0]
frames = np.concatenate([base] * 4)                       # pageable
pinned = torch.from_numpy(frames).pin_memory().numpy()
det = LaneDetector(max_batch=256)
for name, arr in (("pageable", frames), ("pinned", pinned)):
    for _ in range(2):
        det.detect_batch(arr)
    t0 = time.perf_counter()
    for _ in range(5):
        det.detect_batch(arr)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {256 / dt:.0f} frames/s  ({arr.nbytes / dt / 1e9:.1f} GB/s of frames)", flush=True)
det.close()
