"""Stress of the hysteresis / compaction kernel on the GPU box: many random scenes and noise frames at several sizes,
every edge map and point count compared with cv2 itself, each batch run three times (schedule independence)."""
import os, sys
import numpy as np, cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, _native
from oracle.cv2_pipeline import Cv2LaneOracle

def scene(rng, h, w, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = rng.integers(20, 200) + rng.integers(-60, 60) * yy / h + rng.integers(-60, 60) * xx / w
    img = np.clip(np.stack([base + rng.integers(-20, 20) for _ in range(3)], -1), 0, 255).astype(np.uint8)
    for _ in range(int(rng.integers(10, 80))):
        col = tuple(int(v) for v in rng.integers(0, 256, 3))
        p = rng.integers(-20, max(h, w) + 20, 4)
        k = rng.integers(0, 3)
        if k == 0: cv2.line(img, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), col, int(rng.integers(1, 8)))
        elif k == 1: cv2.rectangle(img, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), col, int(rng.choice([-1, 1, 2])))
        else: cv2.circle(img, (int(p[0]) % w, int(p[1]) % h), int(rng.integers(3, 120)), col, int(rng.choice([-1, 1, 3])))
    if kind == 2:   # low-contrast texture: long weak chains
        img = np.clip(img.astype(np.int16) + rng.integers(-14, 15, img.shape), 0, 255).astype(np.uint8)
        img = cv2.GaussianBlur(img, (0, 0), 1.5)
    return img

bad = 0
ref = Cv2LaneOracle()
for seed in range(36):
    rng = np.random.default_rng(5000 + seed)
    h, w = [(1080, 1920), (720, 1280), (480, 640), (2160, 3840), (300, 416), (97, 160)][seed % 6]
    n = 2 if h >= 2000 else 4
    frames = np.stack([scene(rng, h, w, (seed + i) % 3) for i in range(n)])
    det = LaneDetector(max_batch=n, max_segments=4096, debug=True)
    want = [ref.edges(ref.blurred(f)) for f in frames]
    for rep in range(3):
        det.detect_batch(frames)
        for i in range(n):
            got = det._ctx.tap(_native.TAP_EDGES, i)
            if not np.array_equal(got, want[i]):
                bad += 1
                print("MISMATCH seed", seed, "frame", i, "rep", rep, h, w, int((got != want[i]).sum()), flush=True)
            if det.last_records[i]["n_edges"] != int((want[i] != 0).sum()):
                bad += 1
                print("COUNT MISMATCH", seed, i, rep, flush=True)
    det.close()
print("stress done, mismatches:", bad)
