"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``cv2.resize(frame, target_size)`` (INTER_LINEAR, uint8).

SURVEY.md section 8(f) rank 1: the step before the lane path is ``VideoDataLoader.read_frame`` /
``read_frame_at`` (/root/reference/data/loaders/video_loader.py:96-131), whose only pixel work is
``cv2.resize(frame, self.target_size)`` (:108, :128) with OpenCV's default bilinear interpolation.

The arithmetic lives in OpenCV (opencv-python >= 4.5.0, installed 4.13.0.92; source not on this box), so the
restatement below follows the published algorithm (imgproc resize, 8-bit bilinear: 11-bit fixed-point taps,
``HResizeLinear`` then ``VResizeLinear``) and is pinned by tests/test_oracle_resize.py against cv2 itself on
random images, up- and down-scaling, odd sizes, 1 and 3 channels.

    scale = 1 / (dst / src)                                  (double)
    f = float32((d + 0.5) * scale - 0.5); s = floor(f); f -= s
    x: s < 0 -> (s, f) = (0, 0);  s >= src-1 -> (s, f) = (src-1, 0);  taps (s, min(s+1, src-1))
    y: f is NOT clamped, the two rows are clamped to [0, src-1] individually
    a0 = round_half_even((1 - f) * 2048), a1 = round_half_even(f * 2048)          (float32 products)
    H(row, x) = S[row][x0] * a0 + S[row][x1] * a1
    out = sat_u8((((b0 * (H(y0) >> 4)) >> 16) + ((b1 * (H(y1) >> 4)) >> 16) + 2) >> 2)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from typing import Tuple

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _frac(dn: int, sn: int):
    scale = 1.0 / (float(dn) / float(sn))
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _coef(f: np.ndarray):
    a0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int64)
    a1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int64)
    return a0, a1


def x_taps(dn: int, sn: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(x0, x1, a0, a1) per destination column (int64)."""
    s, f = _frac(dn, sn)
    lo = s < 0
    f[lo] = 0
    s[lo] = 0
    hi = s >= sn - 1
    f[hi] = 0
    s[hi] = sn - 1
    a0, a1 = _coef(f)
    return s, np.minimum(s + 1, sn - 1), a0, a1


def y_taps(dn: int, sn: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(y0, y1, b0, b1) per destination row: weights from the unclamped fraction, rows clamped one by one."""
    s, f = _frac(dn, sn)
    b0, b1 = _coef(f)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), b0, b1


def resize_linear(src: np.ndarray, dsize: Tuple[int, int]) -> np.ndarray:
    """``cv2.resize(src, dsize)`` for uint8 ``[H, W]`` / ``[H, W, C]``; ``dsize`` = (width, height)."""
    dw, dh = int(dsize[0]), int(dsize[1])
    src = np.ascontiguousarray(src, dtype=np.uint8)
    sh, sw = src.shape[:2]
    x0, x1, a0, a1 = x_taps(dw, sw)
    y0, y1, b0, b1 = y_taps(dh, sh)
    s = src.astype(np.int64).reshape(sh, sw, -1)
    h = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    r0, r1 = h[y0], h[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])


def resize_batch(frames: np.ndarray, dsize: Tuple[int, int]) -> np.ndarray:
    return np.stack([resize_linear(f, dsize) for f in frames])
