"""Road-layout cues of the reference's SceneClassifier, batched on the GPU.

Mirror of the pixel work in ``/root/reference/src/tagging/scene_classifier.py:145-161``
(``_classify_road_type``): ``gray = cvtColor(frame)``, ``edges = cv2.Canny(gray, 50, 150)`` (no blur,
fixed thresholds), the edge density of the central third of the frame, and
``cv2.HoughLinesP(edges, 1, pi/180, 100, minLineLength=100, maxLineGap=10)`` over the whole frame.
It reuses the lane path's kernels (K1 gray-only mode, K2a/K2b, K4) with other parameters; results are
bit-exact against cv2 (tests/test_gpu_parity.py::test_road_layout_cues_match_cv2).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .. import _native


@dataclass
class RoadLayoutCues:
    center_density: float        # np.sum(center_region > 0) / center_region.size   (:149-150)
    lines: np.ndarray            # int32 [L,4] HoughLinesP segments (x1,y1,x2,y2)    (:156)
    avg_length: float            # mean segment length, 0.0 when there are no lines  (:159-160)

    @property
    def looks_like_intersection(self) -> bool:   # :152
        return self.center_density > 0.15

    @property
    def looks_like_highway(self) -> bool:         # :157-161
        return len(self.lines) > 5 and self.avg_length > 150


class RoadLayoutAnalyzer:
    def __init__(self, *, device: Optional[int] = None, max_batch: int = 64, max_segments: int = 1024):
        self._device, self._max_batch, self._max_segments = device, int(max_batch), int(max_segments)
        self._ctx: Optional[_native.LaneContext] = None
        self._key = None

    def _context(self, h: int, w: int, n: int) -> _native.LaneContext:
        key = (h, w)
        if self._ctx is None or self._key != key or self._ctx.max_batch < min(n, self._max_batch):
            if self._ctx is not None:
                self._ctx.close()
            dev = self._device
            if dev is None:
                import torch
                dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
            ctx = _native.LaneContext(h, w, max(min(n, self._max_batch), 1), np.full((h, w), 255, np.uint8),
                                      device=int(dev), max_segments=self._max_segments, debug=False)
            ctx.set_preprocess(False)                                              # Canny on the plain gray plane
            ctx.set_threshold_lut(np.full(511, 50, np.uint8), np.full(511, 150, np.uint8))   # cv2.Canny(gray, 50, 150)
            ctx.set_hough_params(100, 100, 10)                                     # HoughLinesP(.., 100, 100, 10)
            self._ctx, self._key = ctx, key
        return self._ctx

    def analyze_batch(self, frames: np.ndarray) -> List[RoadLayoutCues]:
        """frames uint8 [N,H,W,3] BGR (host).  One entry per frame."""
        if frames.ndim != 4 or frames.shape[-1] != 3 or frames.dtype != np.uint8:
            raise ValueError("frames must be uint8 [N,H,W,3]")
        n, h, w = frames.shape[:3]
        ctx = self._context(h, w, n)
        out: List[RoadLayoutCues] = []
        pf, pv = np.zeros((1, 2, 3)), np.zeros((1, 2), np.uint8)
        frames = np.ascontiguousarray(frames)
        for a in range(0, n, ctx.max_batch):
            b = min(a + ctx.max_batch, n)
            recs = ctx.detect(frames[a:b], b - a, False, None, 1, pf, pv, 0.7, 1 - 0.7)
            if recs["flags"].any():
                # cv2.HoughLinesP has no segment cap: grow the context to what was found and redo this chunk
                self._max_segments = 1 << int(recs["n_segments_found"].max()).bit_length()
                self._ctx.close()
                self._ctx = None
                ctx = self._context(h, w, n)
                recs = ctx.detect(frames[a:b], b - a, False, None, 1, pf, pv, 0.7, 1 - 0.7)
            # centre-region edge density (:148-150): a popcount over the edge bit-planes on the device
            y0, y1, x0, x1 = h // 3, 2 * h // 3, w // 3, 2 * w // 3
            area = (y1 - y0) * (x1 - x0)
            center = ctx.edge_count_rect(b - a, x0, y0, x1, y1) if area > 0 else np.zeros(b - a, np.int32)
            for i in range(b - a):
                lines = ctx.tap(_native.TAP_SEGMENTS, i)
                avg = float(np.mean(np.sqrt((lines[:, 2] - lines[:, 0]) ** 2.0 + (lines[:, 3] - lines[:, 1]) ** 2.0))) \
                    if len(lines) else 0.0
                out.append(RoadLayoutCues(float(center[i] / area) if area > 0 else float("nan"), lines, avg))
        return out

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
