"""Secondary BASELINE configs (not the bench.py headline line): run on the GPU box, prints one JSON per config.

  config 1  300 x 640x480 generator frames, one stream (the demo's shape): GPU fps (host frames) vs CPU oracle port
  config 2  (add-on) batched standard Hough (K3) device time for the 256 x 1080p bench batch
  config 4  128 x 3840x2160: device-resident fps, per-stage ms, hysteresis rounds
  config 5  batch = 1 streaming latency at 1280x720: p50/p99 per-frame latency vs the CPU path
  noise     64 x 1080p uniform-noise frames (dense edges: worst case for K2/K4), device-resident
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, SyntheticDataGenerator, multi_camera_batch
from oracle.cv2_pipeline import Cv2LaneOracle


def stage_line(det, frames_dev, n, reps=5):
    ctx = det._context(frames_dev.shape[1], frames_dev.shape[2], n)
    pf, pv = np.zeros((1, 2, 3)), np.zeros((1, 2), np.uint8)
    for _ in range(3):
        ctx.detect(frames_dev.data_ptr(), n, True, None, 1, pf, pv, 0.7, 1 - 0.7)
    ctx.set_profiling(True)
    tot, t0 = {}, time.perf_counter()
    for _ in range(reps):
        recs = ctx.detect(frames_dev.data_ptr(), n, True, None, 1, pf, pv, 0.7, 1 - 0.7)
        ms, _ = ctx.stage_ms()
        for k, v in ms.items():
            tot[k] = tot.get(k, 0.0) + v / reps
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ctx.set_profiling(False)
    return tot, wall, recs


def cpu_fps(frames, n=None):
    det = Cv2LaneOracle()
    n = n or len(frames)
    for f in frames[:3]:
        det.detect(f)
    det.reset()
    t0 = time.perf_counter()
    for i in range(n):
        det.detect(frames[i % len(frames)])
    return n / (time.perf_counter() - t0)


def main():
    out = []
    # ---- config 1
    frames = SyntheticDataGenerator().generate_batch(300)
    det = LaneDetector(max_batch=300)
    det.detect_batch(frames); det.reset()
    t0 = time.perf_counter()
    for _ in range(5):
        det.reset(); lanes = det.detect_batch(frames)
    gpu = 5 * 300 / (time.perf_counter() - t0)
    out.append({"config": "1: 300 x 640x480 one stream, host frames through LaneDetector.detect_batch", "gpu_fps": gpu,
                "cpu_fps_1proc": cpu_fps(list(frames), 300), "lanes_found": sum(a is not None and b is not None for a, b in lanes)})
    det.close()
    # ---- config 2 add-on: batched standard Hough (K3) over the bench batch, peaks only
    n = 256
    host = multi_camera_batch(1, n, 1920, 1080, period=64)[0]
    dev = torch.from_numpy(host).cuda()
    det = LaneDetector(max_batch=n)
    det.detect_batch(dev)
    hb = [det._ctx.hough_lines_batch(n, threshold=50, max_peaks=256)[3] for _ in range(4)]
    peaks, counts, _, _ = det._ctx.hough_lines_batch(n, threshold=50, max_peaks=256)
    out.append({"config": "2 (add-on): standard Hough of 256 x 1920x1080 frames, threshold 50, peaks ordered on the device",
                "hough_std_ms": float(min(hb[1:])), "peaks_mean": float(counts.mean())})
    det.close(); del dev
    # ---- config 4
    n = 128
    host = multi_camera_batch(1, n, 3840, 2160, period=4)[0]
    dev = torch.from_numpy(host).cuda()
    det = LaneDetector(max_batch=n)
    tot, wall, recs = stage_line(det, dev, n, reps=3)
    for _ in det.detect_batches([dev] * 2):
        pass
    torch.cuda.synchronize()
    tp = time.perf_counter()
    for _ in det.detect_batches([dev] * 6):          # two batches in flight, back half on the second stream
        pass
    torch.cuda.synchronize()
    piped = 6 * n / (time.perf_counter() - tp)
    out.append({"config": "4: 128 x 3840x2160 device-resident", "fps": n / wall, "fps_pipelined_public_api": piped, "stage_ms": tot,
                "hysteresis_rounds_max": int(recs["hysteresis_rounds"].max()), "segments_mean": float(recs["n_segments"].mean()),
                "k1_frac_of_measured_hbm": n * 4 * 3840 * 2160 / (tot["blur_hist"] / 1e3) / 1e9 / 6556.2,
                "cpu_fps_1proc": cpu_fps(list(host[:4]), 8)})
    det.close(); del dev
    # ---- config 5
    g = SyntheticDataGenerator(1280, 720)
    frames = g.generate_batch(64)
    det = LaneDetector(max_batch=1)
    for f in frames[:8]:
        det.detect(f)
    lat = []
    for i in range(1000):
        t0 = time.perf_counter_ns()
        det.detect(frames[i % 64])
        lat.append((time.perf_counter_ns() - t0) / 1e6)
    ref = Cv2LaneOracle()
    cl = []
    for i in range(300):
        t0 = time.perf_counter_ns()
        ref.detect(frames[i % 64])
        cl.append((time.perf_counter_ns() - t0) / 1e6)
    out.append({"config": "5: batch=1 1280x720 streaming latency through LaneDetector.detect (host frame in, LaneLines out)",
                "gpu_ms_p50": float(np.percentile(lat, 50)), "gpu_ms_p99": float(np.percentile(lat, 99)),
                "cpu_ms_p50": float(np.percentile(cl, 50)), "cpu_ms_p99": float(np.percentile(cl, 99))})
    det.close()
    # ---- dense noise
    n = 64
    rng = np.random.default_rng(0)
    host = rng.integers(0, 256, (n, 1080, 1920, 3), dtype=np.uint8)
    dev = torch.from_numpy(host).cuda()
    det = LaneDetector(max_batch=n, max_segments=4096)
    tot, wall, recs = stage_line(det, dev, n, reps=2)
    out.append({"config": "noise: 64 x 1080p uniform noise device-resident (dense worst case)", "fps": n / wall, "stage_ms": tot,
                "roi_points_mean": float(recs["n_roi_points"].mean()), "hysteresis_rounds_max": int(recs["hysteresis_rounds"].max()),
                "segments_mean": float(recs["n_segments"].mean())})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
