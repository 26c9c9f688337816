"""GPU-box probe: dense worst case (uniform-noise 1080p frames, ~88 k ROI edge points each): stage times, for A/B of PPHT choices."""
import json, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device="cuda").manual_seed(1)
dev = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
det = LaneDetector(max_batch=n, max_segments=4096)
ctx = det._context(1080, 1920, n)
ctx.close(); det._ctx = None
from multimodal_autonomous_driving_perception_and_planning_b200 import _native
ctx = _native.LaneContext(1080, 1920, n, det._get_roi_mask((1080, 1920)), device=0, max_segments=4096)
pf, pv = np.zeros((1, 2, 3)), np.zeros((1, 2), np.uint8)
ctx.detect(dev.data_ptr(), n, True, None, 1, pf, pv, 0.7, 0.3)
ctx.set_profiling(True)
recs = ctx.detect(dev.data_ptr(), n, True, None, 1, pf, pv, 0.7, 0.3)
ms, _ = ctx.stage_ms()
import hashlib
print(json.dumps({"lib": os.environ.get("LANE_B200_LIB"), "ppht_ms": ms["ppht"], "paths": ctx.last_paths(), "segments_mean": float(recs["n_segments"].mean()),
                  "points_mean": float(recs["n_roi_points"].mean()), "records_sha": hashlib.sha256(recs.tobytes()).hexdigest()[:16]}))
