"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SceneClassifier's image statistics.

Reference lines (/root/reference/src/tagging/scene_classifier.py):
    :183-186  hsv = cv2.cvtColor(frame, cv2.COLOR_BGR2HSV); green = cv2.inRange(hsv, (35,40,40), (85,255,255));
              green_ratio = np.sum(green > 0) / green.size
    :237-238  gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY); avg_brightness = np.mean(gray)
    :254      laplacian_var = cv2.Laplacian(gray, cv2.CV_64F).var()

The arithmetic lives in OpenCV (opencv-python >= 4.5.0, installed 4.13.0.92; source not on this box).  8-bit BGR2HSV
is integer arithmetic with two 12-bit fixed-point division tables; Laplacian with the default ksize = 1 is the
3x3 kernel [0 1 0; 1 -4 1; 0 1 0] under BORDER_REFLECT_101.  Pinned against cv2 in tests/test_oracle_scene_stats.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from dataclasses import dataclass

import numpy as np


def hsv_tables():
    sdiv = np.zeros(256, np.int64)
    hdiv = np.zeros(256, np.int64)
    for i in range(1, 256):
        sdiv[i] = int(np.rint((255 << 12) / (1.0 * i)))
        hdiv[i] = int(np.rint((180 << 12) / (6.0 * i)))
    return sdiv, hdiv


def bgr2hsv(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_BGR2HSV) for uint8 (H in 0..179)."""
    sdiv, hdiv = hsv_tables()
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    s = (diff * sdiv[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hdiv[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def gray_q15(img: np.ndarray) -> np.ndarray:
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def laplacian(gray: np.ndarray) -> np.ndarray:
    """cv2.Laplacian(gray, cv2.CV_64F) as exact integers (int64)."""
    h, w = gray.shape
    g = gray.astype(np.int64)
    yi = np.arange(-1, h + 1)
    xi = np.arange(-1, w + 1)
    fold = lambda p, n: np.zeros_like(p) if n == 1 else np.where(p < 0, -p, np.where(p >= n, 2 * n - 2 - p, p))
    gp = g[fold(yi, h)][:, fold(xi, w)]
    return gp[:-2, 1:-1] + gp[2:, 1:-1] + gp[1:-1, :-2] + gp[1:-1, 2:] - 4 * gp[1:-1, 1:-1]


@dataclass
class FrameSums:
    sum_gray: int
    sum_laplacian: int
    sum_laplacian_sq: int
    green_pixels: int
    n_pixels: int


def frame_sums(frame: np.ndarray) -> FrameSums:
    gray = gray_q15(frame)
    lap = laplacian(gray)
    hsv = bgr2hsv(frame)
    green = (hsv[..., 0] >= 35) & (hsv[..., 0] <= 85) & (hsv[..., 1] >= 40) & (hsv[..., 2] >= 40)
    return FrameSums(int(gray.astype(np.int64).sum()), int(lap.sum()), int((lap * lap).sum()), int(green.sum()),
                     int(gray.size))


def cues_from_sums(s: FrameSums):
    """(avg_brightness, laplacian_var, green_ratio) as float64 from the exact sums."""
    n = s.n_pixels
    return s.sum_gray / n, (s.sum_laplacian_sq * n - s.sum_laplacian * s.sum_laplacian) / (n * n), s.green_pixels / n
