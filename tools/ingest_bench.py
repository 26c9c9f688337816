"""Frame-ingest throughput on the GPU box: batched cv2.resize on the device (K0) vs cv2.resize on the host cores.
Device-resident source and destination; CUDA events; algorithmic bytes = touched source bytes + destination bytes."""
import json, os, sys, time
import numpy as np, cv2, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_autonomous_driving_perception_and_planning_b200 import FrameIngest, multi_camera_batch

peak = 6556.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for (sw, sh, dw, dh, n) in [(1920, 1080, 640, 480, 256), (1920, 1080, 1280, 720, 256), (3840, 2160, 1920, 1080, 64)]:
    base = multi_camera_batch(1, 8, sw, sh)[0]
    host = np.concatenate([base] * (n // 8))
    src = torch.from_numpy(host).cuda()
    ing = FrameIngest((dw, dh))
    for _ in range(3):
        out = ing.resize_batch(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = ing.resize_batch(src)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # source bytes actually touched: two rows per destination row, at most all columns
    rows_touched = min(sh, 2 * dh)
    cols_touched = min(sw, 2 * dw)
    algo = n * (rows_touched * cols_touched * 3 + dh * dw * 3)
    t0 = time.perf_counter()
    for f in host[:32]:
        cv2.resize(f, (dw, dh))
    cpu_fps = 32 / (time.perf_counter() - t0)
    ok = bool(np.array_equal(out[:2].cpu().numpy(), np.stack([cv2.resize(f, (dw, dh)) for f in host[:2]])))
    print(json.dumps({"ingest": f"{n} x {sw}x{sh} -> {dw}x{dh}", "ms": ms, "frames_per_s": n / (ms / 1e3),
                      "algorithmic_GBps": algo / (ms / 1e3) / 1e9, "frac_of_measured_hbm": algo / (ms / 1e3) / 1e9 / peak,
                      "cpu_fps_1thread_default_cv2": cpu_fps, "bit_exact_vs_cv2": ok}), flush=True)

# ---- scene statistics (K6): one pass over the BGR frames, 3 B/px read, nothing written
from multimodal_autonomous_driving_perception_and_planning_b200.perception.scene_stats import SceneStatsAnalyzer
import ctypes as C
from multimodal_autonomous_driving_perception_and_planning_b200 import _native
for (sw, sh, n) in [(1920, 1080, 256)]:
    base = multi_camera_batch(1, 8, sw, sh)[0]
    src = torch.from_numpy(np.concatenate([base] * (n // 8))).cuda()
    an = SceneStatsAnalyzer()
    for _ in range(2):
        an.analyze_batch(src)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        res = an.analyze_batch(src)          # blocking call (records come back to the host)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    t0 = time.perf_counter()
    for f in base:
        g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY); np.mean(g); cv2.Laplacian(g, cv2.CV_64F).var()
        m = cv2.inRange(cv2.cvtColor(f, cv2.COLOR_BGR2HSV), (35, 40, 40), (85, 255, 255)); np.sum(m > 0) / m.size
    cpu_fps = len(base) / (time.perf_counter() - t0)
    algo = n * sh * sw * 3
    print(json.dumps({"scene_stats": f"{n} x {sw}x{sh}", "ms_wall_per_call": ms, "frames_per_s": n / (ms / 1e3),
                      "algorithmic_GBps": algo / (ms / 1e3) / 1e9, "frac_of_measured_hbm": algo / (ms / 1e3) / 1e9 / peak,
                      "cpu_fps_default_cv2": cpu_fps, "avg_brightness": res[0].avg_brightness,
                      "laplacian_var": res[0].laplacian_var, "green_ratio": res[0].green_ratio}), flush=True)
