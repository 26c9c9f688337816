"""numpy (pageable host) input through the context-free entry points: scene statistics and resize."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_autonomous_driving_perception_and_planning_b200 import FrameIngest, multi_camera_batch
from multimodal_autonomous_driving_perception_and_planning_b200.perception.scene_stats import SceneStatsAnalyzer
frames = np.concatenate([multi_camera_batch(1, 32, 1920, 1080)[0]] * 4)       # 128 x 1080p, 796 MB, pageable
an, ing = SceneStatsAnalyzer(), FrameIngest((640, 480))
for name, fn in (("scene_stats", lambda: an.analyze_batch(frames)), ("resize->480p", lambda: ing.resize_batch(frames))):
    fn()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {len(frames) / dt:.0f} frames/s ({frames.nbytes / dt / 1e9:.1f} GB/s of input)", flush=True)
