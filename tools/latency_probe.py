"""batch = 1 latency of LaneDetector.detect at 720p / 1080p (host frame in, LaneLines out), p50 / p99 over 400 frames."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, multi_camera_batch
for (w, h) in [(1280, 720), (1920, 1080)]:
    frames = multi_camera_batch(1, 64, w, h)[0]
    det = LaneDetector(max_batch=1)
    for i in range(20):
        det.detect(frames[i % 64])
    ts = []
    for i in range(400):
        t0 = time.perf_counter()
        det.detect(frames[i % 64])
        ts.append((time.perf_counter() - t0) * 1e3)
    ms, ln = det._ctx.stage_ms()
    print(f"{w}x{h}: p50 {np.percentile(ts, 50):.3f} ms  p99 {np.percentile(ts, 99):.3f} ms", flush=True)
    det.close()
