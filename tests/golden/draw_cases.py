"""Inputs of the drawing goldens, shared by make_golden_draw.py (which runs the unmodified reference on them) and the
tests (which run the oracle, the host expansion and the GPU kernel on them)."""
import numpy as np

STREAMS = [(640, 480, 12), (1920, 1080, 3)]          # (width, height, frames) of generator streams, detector run in order
N_RANDOM = 48


def random_case(seed):
    """A noise frame with two random quadratic lanes (points may leave the image on every side), or a missing side."""
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(40, 360)), int(rng.integers(40, 480))
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    pts = np.zeros((2, 50, 2), np.int32)
    valid = np.zeros(2, np.uint8)
    for s in range(2):
        if rng.random() < 0.15:
            continue
        c = np.array([rng.normal(0, 2e-3), rng.normal(0, 1.0), rng.normal(w / 2, w / 2)])
        y = np.linspace(0.6 * h, h, 50)
        pts[s] = np.column_stack([np.polyval(c, y), y]).astype(np.int32)
        valid[s] = 1
    offset = None if seed % 7 == 0 else float(rng.integers(-300, 300)) / 2
    return frame, pts, valid, offset
