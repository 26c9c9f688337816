"""Stress of the device rasterisers on the GPU box: thousands of random cv2 call mixes (lines of every thickness, filled /
outlined rectangles, circles, polygons with up to 40 vertices, weighted polygon fills, polylines, row runs; vertices up to
twice the frame size outside it) at several frame sizes, many frames with different lists per launch, compared with cv2
itself; and random lane overlays through k7_lanes and through the primitive lists.  Prints one summary line."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from draw_util import cv2_draw_lanes
from multimodal_autonomous_driving_perception_and_planning_b200 import DrawList
from multimodal_autonomous_driving_perception_and_planning_b200.visualization.overlays import draw_lanes_arrays

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 2024)
bad = frames_checked = calls = 0
for rep in range(60):
    h, w = [(97, 131), (480, 640), (1080, 1920), (33, 500), (720, 1280), (256, 256)][rep % 6]
    n = 4 if h >= 700 else 24
    start = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    ref = start.copy()
    dl = DrawList(n)

    def rp():
        return (int(rng.integers(-w, 2 * w)), int(rng.integers(-h, 2 * h)))

    for f in range(n):
        for _ in range(int(rng.integers(1, 12))):
            col = tuple(int(c) for c in rng.integers(0, 256, 3))
            kind = int(rng.integers(0, 7))
            calls += 1
            if kind == 0:
                a, b, th = rp(), rp(), int(rng.integers(1, 12))
                cv2.line(ref[f], a, b, col, th); dl.line(f, a, b, col, th)
            elif kind == 1:
                a, b, th = rp(), rp(), int(rng.choice([-1, 1, 2, 5]))
                cv2.rectangle(ref[f], a, b, col, th); dl.rectangle(f, a, b, col, th)
            elif kind == 2:
                c, r = rp(), int(rng.integers(0, 200))
                cv2.circle(ref[f], c, r, col, -1); dl.circle(f, c, r, col, -1)
            elif kind == 3:
                pts = np.array([rp() for _ in range(int(rng.integers(1, 40)))], np.int32)
                cv2.fillPoly(ref[f], [pts], col); dl.fillPoly(f, pts, col)
            elif kind == 4:
                pts = np.array([rp() for _ in range(int(rng.integers(1, 20)))], np.int32)
                th, closed = int(rng.integers(1, 7)), bool(rng.integers(0, 2))
                cv2.polylines(ref[f], [pts], closed, col, th); dl.polylines(f, pts, closed, col, th)
            elif kind == 5:
                pts = np.array([rp() for _ in range(int(rng.integers(3, 30)))], np.int32)
                al, be = [(0.7, 0.3), (0.5, 0.5), (0.25, 0.6)][int(rng.integers(0, 3))]
                o = ref[f].copy(); cv2.fillPoly(o, [pts], col)
                ref[f] = cv2.addWeighted(ref[f], al, o, be, 0); dl.fillPoly_weighted(f, pts, col, al, be)
            else:
                y0, cnt = int(rng.integers(-5, h)), int(rng.integers(1, 40))
                x1, x2 = int(rng.integers(-10, w + 10)), int(rng.integers(-10, w + 10))
                cols = [tuple(int(c) for c in rng.integers(0, 256, 3)) for _ in range(cnt)]
                for i, c in enumerate(cols):
                    cv2.line(ref[f], (x1, y0 + i), (x2, y0 + i), c, 1)
                dl.rows(f, y0, x1, x2, cols)
    got = start.copy()
    dl.execute(got)
    bad += int((got != ref).any(axis=(1, 2, 3)).sum())
    frames_checked += n
lane_bad = lane_frames = 0
for rep in range(30):
    h, w = [(480, 640), (1080, 1920), (173, 301)][rep % 3]
    n = 8 if h >= 1000 else 32
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    y = np.linspace(0.6 * h, h, 50)
    pts = np.zeros((n, 2, 50, 2), np.int32)
    valid = (rng.random((n, 2)) > 0.1).astype(np.uint8)
    for f in range(n):
        for s in range(2):
            c = [rng.normal(0, 3e-3), rng.normal(0, 1.5), rng.normal(w / 2, w)]
            pts[f, s] = np.column_stack([np.polyval(c, y), y]).astype(np.int32)
    for fill in (True, False):
        got = frames.copy()
        draw_lanes_arrays(got, pts[:, 0], valid[:, 0], pts[:, 1], valid[:, 1], fill)
        for f in range(n):
            ref = cv2_draw_lanes(frames[f].copy(), pts[f, 0] if valid[f, 0] else None, pts[f, 1] if valid[f, 1] else None, fill)
            lane_bad += not np.array_equal(ref, got[f])
            lane_frames += 1
print(f"draw stress done: {calls} cv2 calls on {frames_checked} frames, {bad} frames differ; "
      f"{lane_frames} lane overlays ({os.environ.get('LANE_B200_DRAW_LANES', 'k7_lanes')}), {lane_bad} differ")
