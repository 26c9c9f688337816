"""Stress of the PPHT kernel on the GPU box: random scenes at several sizes with random HoughLinesP parameters,
segments compared with cv2.HoughLinesP on the same masked edge map, each batch run twice."""
import os, sys
import numpy as np, cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, _native
from oracle.cv2_pipeline import Cv2LaneOracle

def scene(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    base = rng.integers(20, 200) + rng.integers(-60, 60) * yy / h + rng.integers(-60, 60) * xx / w
    img = np.clip(np.stack([base + rng.integers(-20, 20) for _ in range(3)], -1), 0, 255).astype(np.uint8)
    for _ in range(int(rng.integers(5, 60))):
        col = tuple(int(v) for v in rng.integers(0, 256, 3))
        p = rng.integers(-20, max(h, w) + 20, 4)
        k = rng.integers(0, 3)
        if k == 0: cv2.line(img, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), col, int(rng.integers(1, 8)))
        elif k == 1: cv2.rectangle(img, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), col, int(rng.choice([-1, 1, 2])))
        else: cv2.circle(img, (int(p[0]) % w, int(p[1]) % h), int(rng.integers(3, 120)), col, int(rng.choice([-1, 1, 3])))
    return img

bad = total = 0
ref = Cv2LaneOracle()
for seed in range(40):
    rng = np.random.default_rng(9000 + seed)
    h, w = [(1080, 1920), (720, 1280), (480, 640), (2160, 3840), (300, 416)][seed % 5]
    n = 2 if h >= 2000 else 3
    thr, min_len, gap = (50, 50, 150) if seed % 3 == 0 else (int(rng.integers(10, 80)), int(rng.integers(5, 120)), int(rng.integers(0, 200)))
    frames = np.stack([scene(rng, h, w) for _ in range(n)])
    det = LaneDetector(max_batch=n, max_segments=4096, debug=True)
    det._context(h, w, n).set_hough_params(thr, min_len, gap)
    want = []
    for f in frames:
        masked = ref.masked(ref.edges(ref.blurred(f)))
        lines = cv2.HoughLinesP(masked, rho=1, theta=np.pi / 180, threshold=thr, minLineLength=min_len, maxLineGap=gap)
        want.append(np.zeros((0, 4), np.int32) if lines is None else lines.reshape(-1, 4))
    for rep in range(2):
        det.detect_batch(frames)
        for i in range(n):
            got = det._ctx.tap(_native.TAP_SEGMENTS, i)
            total += 1
            if not np.array_equal(got, want[i]):
                bad += 1
                print("MISMATCH seed", seed, "frame", i, "rep", rep, (h, w), (thr, min_len, gap), len(got), len(want[i]), flush=True)
    det.close()
print("ppht stress done:", total, "comparisons,", bad, "mismatches; segments per frame up to", max(len(x) for x in want))
