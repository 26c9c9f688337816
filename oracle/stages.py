"""Stage-by-stage CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Every function cites the reference line whose behaviour it restates; paths are under
/root/reference/.  Pixel stages call the C restatement in lane_oracle.c; the fit tail is
NumPy because the reference itself calls numpy there.
"""
from __future__ import annotations

import ctypes
import math
import warnings
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import lib

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i16p = ctypes.POINTER(ctypes.c_int16)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def _p(a: Optional[np.ndarray], t):
    return a.ctypes.data_as(t) if a is not None else None


def gray(frame: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, BGR2GRAY) -- src/perception/lane_detector.py:69."""
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    h, w = frame.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().orc_gray(_p(frame, _u8p), h, w, _p(out, _u8p))
    return out


def blur5(g: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(gray, (5, 5), 0) -- src/perception/lane_detector.py:72."""
    g = np.ascontiguousarray(g, dtype=np.uint8)
    h, w = g.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_blur5(_p(g, _u8p), h, w, _p(out, _u8p))
    return out


def hist256(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros(256, np.uint32)
    lib().orc_hist(_p(img, _u8p), ctypes.c_long(img.size), _p(out, _u32p))
    return out


def median_x2(hist: np.ndarray, n: int) -> int:
    """2 * np.median(blurred) from a 256-bin histogram -- src/perception/lane_detector.py:79."""
    hist = np.ascontiguousarray(hist, dtype=np.uint32)
    return int(lib().orc_median_x2(_p(hist, _u32p), ctypes.c_long(n)))


def threshold_lut() -> Tuple[np.ndarray, np.ndarray]:
    """low/high for every possible 2*median, by evaluating the reference's own float64
    expressions -- src/perception/lane_detector.py:80-81."""
    low = np.empty(511, np.uint8)
    high = np.empty(511, np.uint8)
    for k in range(511):
        m = np.float64(k) / 2.0  # np.median returns float64
        low[k] = int(max(0, 0.7 * m))
        high[k] = int(min(255, 1.3 * m))
    return low, high


def thresholds(blurred: np.ndarray) -> Tuple[int, int, int]:
    """(low, high, 2*median) of a blurred frame."""
    lo, hi = threshold_lut()
    m2 = median_x2(hist256(blurred), blurred.size)
    return int(lo[m2]), int(hi[m2]), m2


@dataclass
class CannyTaps:
    dx: np.ndarray
    dy: np.ndarray
    mag: np.ndarray
    cls: np.ndarray  # 0 none, 1 weak candidate, 2 strong (before hysteresis)
    edges: np.ndarray


def canny(blurred: np.ndarray, low: int, high: int) -> CannyTaps:
    """cv2.Canny(blurred, low, high) -- src/perception/lane_detector.py:83."""
    b = np.ascontiguousarray(blurred, dtype=np.uint8)
    h, w = b.shape
    dx = np.empty((h, w), np.int16)
    dy = np.empty((h, w), np.int16)
    mag = np.empty((h, w), np.uint16)
    cls = np.empty((h, w), np.uint8)
    edges = np.empty((h, w), np.uint8)
    lib().orc_canny(_p(b, _u8p), h, w, int(low), int(high), _p(dx, _i16p), _p(dy, _i16p),
                    _p(mag, _u16p), _p(cls, _u8p), _p(edges, _u8p))
    return CannyTaps(dx, dy, mag, cls, edges)


def default_roi_vertices(h: int, w: int) -> np.ndarray:
    """Default trapezoid -- src/perception/lane_detector.py:55-60."""
    return np.array([[(int(w * 0.1), h), (int(w * 0.4), int(h * 0.6)),
                      (int(w * 0.6), int(h * 0.6)), (int(w * 0.9), h)]], dtype=np.int32)


def roi_mask(h: int, w: int, roi_vertices: Optional[np.ndarray] = None) -> np.ndarray:
    """cv2.fillPoly mask -- src/perception/lane_detector.py:47-64.  The rasteriser is cv2's
    own (the mask is frame independent, so product and oracle both take it from cv2)."""
    import cv2
    verts = roi_vertices if roi_vertices is not None else default_roi_vertices(h, w)
    mask = np.zeros((h, w), np.uint8)
    cv2.fillPoly(mask, verts, 255)
    return mask


def hough_numrho(h: int, w: int) -> int:
    return 2 * (w + h) + 1


def hough_accum(img: np.ndarray) -> np.ndarray:
    """Padded accumulator [182][numrho+2] that cv2.HoughLines(img, 1, pi/180, thr) votes into."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    na = int(lib().orc_hough_numangle())
    acc = np.zeros((na + 2, hough_numrho(h, w) + 2), np.int32)
    lib().orc_hough_accum(_p(img, _u8p), h, w, _p(acc, _i32p))
    return acc


def hough_peaks(acc: np.ndarray, h: int, w: int, threshold: int, max_out: int = 1 << 16) -> np.ndarray:
    """Local maxima as rows (r, n, votes) in cv2's output order."""
    acc = np.ascontiguousarray(acc, dtype=np.int32)
    out = np.empty((max_out, 3), np.int32)
    cnt = int(lib().orc_hough_peaks(_p(acc, _i32p), h, w, int(threshold), _p(out, _i32p), max_out))
    return out[:min(cnt, max_out)].copy()


def hough_tables_std() -> Tuple[np.ndarray, np.ndarray]:
    na = int(lib().orc_hough_numangle())
    s = np.empty(na, np.float32)
    c = np.empty(na, np.float32)
    lib().orc_hough_tables_std(_p(s, _f32p), _p(c, _f32p), na)
    return s, c


def houghp_tables() -> Tuple[np.ndarray, np.ndarray]:
    c = np.empty(180, np.float32)
    s = np.empty(180, np.float32)
    lib().orc_houghp_tables(_p(c, _f32p), _p(s, _f32p), 180)
    return c, s


def houghp(img: np.ndarray, threshold: int = 50, min_line_length: int = 50, max_line_gap: int = 150,
           max_lines: int = 4096, with_trace: bool = False):
    """cv2.HoughLinesP(img, 1, pi/180, threshold, minLineLength, maxLineGap) as int32 [L,4]
    -- src/perception/lane_detector.py:94-101."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    lines = np.empty((max_lines, 4), np.int32)
    trace = np.zeros(8, np.int64)
    n = int(lib().orc_houghp(_p(img, _u8p), h, w, int(threshold), int(round(min_line_length)),
                             int(round(max_line_gap)), _p(lines, _i32p), max_lines, _p(trace, _i64p)))
    out = lines[:min(n, max_lines)].copy()
    return (out, trace) if with_trace else out


# --------------------------------------------------------------------------------------
# fit tail (NumPy, as in the reference)
# --------------------------------------------------------------------------------------

def separate(lines: np.ndarray, width: int) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """Slope/side split -- src/perception/lane_detector.py:105-134.  `lines` is int32 [L,4]."""
    left, right = [], []
    cx = width / 2
    for seg in np.asarray(lines, dtype=np.int32).reshape(-1, 4):
        x1, y1, x2, y2 = seg  # numpy int32 scalars, as in the reference
        if x2 == x1:
            continue
        slope = (y2 - y1) / (x2 - x1)
        if abs(slope) < 0.3:
            continue
        mid = (x1 + x2) / 2
        if slope < 0 and mid < cx:
            left.append(seg)
        elif slope > 0 and mid > cx:
            right.append(seg)
    return left, right


@dataclass
class SideFit:
    raw: np.ndarray          # polyfit output before smoothing
    coeffs: np.ndarray       # after EMA
    points: np.ndarray       # int32 [50,2]
    confidence: float
    n_lines: int


def fit_side(segs: List[np.ndarray], height: int, prev: Optional[np.ndarray],
             smoothing: float = 0.7) -> Optional[SideFit]:
    """Quadratic fit + EMA + 50 sample points -- src/perception/lane_detector.py:136-176."""
    if not segs:
        return None
    xs, ys = [], []
    for x1, y1, x2, y2 in segs:
        xs.extend([x1, x2])
        ys.extend([y1, y2])
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # RankWarning for 2 distinct y values
            raw = np.polyfit(ys, xs, 2)
    except np.linalg.LinAlgError:
        return None
    c = raw
    if prev is not None:
        c = smoothing * prev + (1 - smoothing) * raw
    yp = np.linspace(height * 0.6, height, 50)
    xp = np.polyval(c, yp)
    with np.errstate(invalid="ignore"):
        pts = np.column_stack((xp, yp)).astype(np.int32)
    return SideFit(raw=raw, coeffs=c, points=pts, confidence=min(1.0, len(segs) / 10), n_lines=len(segs))


def center_offset(width: int, left: Optional[SideFit], right: Optional[SideFit]) -> Optional[float]:
    """src/perception/lane_detector.py:253-272."""
    if left is None or right is None:
        return None
    lane_center = (left.points[-1, 0] + right.points[-1, 0]) / 2
    return float(width / 2 - lane_center)


@dataclass
class FrameTrace:
    gray: np.ndarray
    blurred: np.ndarray
    median_x2: int
    low: int
    high: int
    canny: CannyTaps
    masked: np.ndarray
    lines: np.ndarray
    left: Optional[SideFit]
    right: Optional[SideFit]
    offset: Optional[float]


@dataclass
class StageOracle:
    """Whole pipeline with every intermediate kept; state mirrors LaneDetector's EMA
    (src/perception/lane_detector.py:43-45, :210-216)."""
    roi_vertices: Optional[np.ndarray] = None
    smoothing_factor: float = 0.7
    prev_left: Optional[np.ndarray] = None
    prev_right: Optional[np.ndarray] = None
    _mask_cache: dict = field(default_factory=dict)

    def reset(self):
        self.prev_left = None
        self.prev_right = None

    def mask(self, h, w):
        key = (h, w)
        if key not in self._mask_cache:
            self._mask_cache[key] = roi_mask(h, w, self.roi_vertices)
        return self._mask_cache[key]

    def step(self, frame: np.ndarray) -> FrameTrace:
        h, w = frame.shape[:2]
        g = gray(frame)
        b = blur5(g)
        low, high, m2 = thresholds(b)
        taps = canny(b, low, high)
        masked = taps.edges & self.mask(h, w)
        lines = houghp(masked)
        ls, rs = separate(lines, w)
        lf = fit_side(ls, h, self.prev_left, self.smoothing_factor)
        rf = fit_side(rs, h, self.prev_right, self.smoothing_factor)
        if lf is not None:
            self.prev_left = lf.coeffs
        if rf is not None:
            self.prev_right = rf.coeffs
        return FrameTrace(g, b, m2, low, high, taps, masked, lines, lf, rf, center_offset(w, lf, rf))
