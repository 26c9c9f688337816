"""Small drawing / streaming cases for compute-sanitizer (memcheck, racecheck): k7_draw with every primitive kind
(mask + blend, polygon fill with many crossings, text blit, thick lines) and the two-stream streaming path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from draw_util import cv2_draw_lanes, cv2_offset_indicator, random_mix
from multimodal_autonomous_driving_perception_and_planning_b200 import (DrawList, LaneDetector, OverlayRenderer,
                                                                        SyntheticDataGenerator, draw_lanes_batch)

rng = np.random.default_rng(0)
bad = 0
for t in range(40):
    img, ref, dl = random_mix(rng, lambda: DrawList(1))
    mine = img.copy()[None]
    dl.execute(mine)
    bad += not np.array_equal(ref, mine[0])
print("random mixes mismatching:", bad)
w, h, n = 320, 240, 4
gen = SyntheticDataGenerator(w, h)
frames = gen.generate_batch_device(n, start_frame=0)
assert np.array_equal(frames.cpu().numpy(), SyntheticDataGenerator(w, h).generate_batch(n, start_frame=0))
det = LaneDetector(max_batch=n)
lanes = det.detect_batch(frames)
out = draw_lanes_batch(frames.clone(), lanes)
offs = [det.get_lane_center_offset(w, l, r) for l, r in lanes]
OverlayRenderer().draw_lane_offset_indicator_batch(out, offs)
host = frames.cpu().numpy()
for i, (l, r) in enumerate(lanes):
    ref = cv2_offset_indicator(cv2_draw_lanes(host[i].copy(), None if l is None else l.points, None if r is None else r.points), offs[i])
    assert np.array_equal(ref, out[i].cpu().numpy()), i
print("detect -> draw ok")
# streaming, two batches in flight (back half on the second stream)
ctx = det._context(h, w, n)
b = [gen.generate_batch_device(n, start_frame=10 * k) for k in range(4)]
pf, pv = np.zeros((1, 2, 3)), np.zeros((1, 2), np.uint8)
ctx.enqueue(b[0].data_ptr(), n, None, 1, pf, pv, 0.7, 1 - 0.7)
for k in range(4):
    if k + 1 < 4:
        ctx.enqueue(b[k + 1].data_ptr(), n, None, 1, None, None, 0.7, 1 - 0.7)
    ctx.collect(pf, pv)
torch.cuda.synchronize()
print("streaming ok")
det.close()
