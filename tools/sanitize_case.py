"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both K1/K2 code paths, PPHT v3 and v2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, SyntheticDataGenerator

frames = SyntheticDataGenerator(640, 480).generate_batch(3)
det = LaneDetector(max_batch=3, debug=True)
lanes = det.detect_batch(frames)
print("480p lanes:", [(a is not None, b is not None) for a, b in lanes], det.last_records["n_segments"])
det.close()
rng = np.random.default_rng(0)
noise = rng.integers(0, 256, (2, 240, 320, 3), dtype=np.uint8)     # dense: PPHT list overflows shared memory -> v2
det = LaneDetector(max_batch=2, max_segments=2048)
det.detect_batch(noise)
print("noise segments:", det.last_records["n_segments"], det.last_records["n_roi_points"])
det.close()
odd = rng.integers(0, 256, (2, 121, 203, 3), dtype=np.uint8)       # generic-width fallback kernels
det = LaneDetector(max_batch=2)
det.detect_batch(odd)
print("odd-size edges:", det.last_records["n_edges"])
det.close()
