"""The reference's call sequence on the real cv2/numpy (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/perception/lane_detector.py:178-218 (`detect`) stage by stage,
calling the very same third-party functions with the same literals, so that on the GPU box
(where /root/reference is not mounted) there is still an executable statement of "what the
reference computes" -- used as the final arbiter in tests and as the CPU baseline in
bench.py (`cpu_baseline.kind = "port"`: our restatement of the call sequence, the reference's
own OpenCV/NumPy binaries underneath).
"""
from __future__ import annotations

from typing import Optional, Tuple

import cv2
import numpy as np

from .stages import SideFit, center_offset, default_roi_vertices, fit_side, separate


class Cv2LaneOracle:
    def __init__(self, roi_vertices: Optional[np.ndarray] = None):
        self.roi_vertices = roi_vertices
        self.smoothing_factor = 0.7
        self.prev_left = None
        self.prev_right = None

    def reset(self):
        self.prev_left = None
        self.prev_right = None

    # -- stages, one per reference line ------------------------------------------------
    @staticmethod
    def blurred(frame):                                   # lane_detector.py:69-72
        return cv2.GaussianBlur(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (5, 5), 0)

    @staticmethod
    def edges(blurred):                                   # lane_detector.py:79-83
        m = np.median(blurred)
        return cv2.Canny(blurred, int(max(0, 0.7 * m)), int(min(255, 1.3 * m)))

    def masked(self, edges):                              # lane_detector.py:47-64, :86-90
        h, w = edges.shape[:2]
        verts = self.roi_vertices if self.roi_vertices is not None else default_roi_vertices(h, w)
        mask = np.zeros((h, w), np.uint8)
        cv2.fillPoly(mask, verts, 255)
        return cv2.bitwise_and(edges, mask)

    @staticmethod
    def segments(masked):                                 # lane_detector.py:94-103
        lines = cv2.HoughLinesP(masked, rho=1, theta=np.pi / 180, threshold=50,
                                minLineLength=50, maxLineGap=150)
        return np.zeros((0, 4), np.int32) if lines is None else lines.reshape(-1, 4)

    def detect(self, frame) -> Tuple[Optional[SideFit], Optional[SideFit]]:
        h, w = frame.shape[:2]
        segs = self.segments(self.masked(self.edges(self.blurred(frame))))
        ls, rs = separate(segs, w)                        # lane_detector.py:105-134
        lf = fit_side(ls, h, self.prev_left, self.smoothing_factor)    # :136-176
        rf = fit_side(rs, h, self.prev_right, self.smoothing_factor)
        if lf is not None:                                # :210-216
            self.prev_left = lf.coeffs
        if rf is not None:
            self.prev_right = rf.coeffs
        return lf, rf

    @staticmethod
    def offset(width, lf, rf):                            # lane_detector.py:253-272
        return center_offset(width, lf, rf)
