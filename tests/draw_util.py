"""Shared helpers of the drawing tests: the reference's cv2 call sequences restated on the real cv2 (the arbiter where
/root/reference is not mounted), golden access, and random cv2-level command mixes."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np

from util import h16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def draw_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "draw_golden.npz"))


def cv2_draw_lanes(frame, left_pts, right_pts, fill_lane=True):
    """/root/reference/src/perception/lane_detector.py:220-251 on cv2 (``*_pts``: int32[50,2] or None)."""
    overlay = frame.copy()
    if fill_lane and left_pts is not None and right_pts is not None:
        pts = np.vstack([left_pts, right_pts[::-1]])
        cv2.fillPoly(overlay, [pts], (0, 255, 100))
        frame = cv2.addWeighted(frame, 0.7, overlay, 0.3, 0)
    if left_pts is not None:
        cv2.polylines(frame, [left_pts], False, (255, 0, 0), 3)
    if right_pts is not None:
        cv2.polylines(frame, [right_pts], False, (0, 0, 255), 3)
    return frame


def cv2_offset_indicator(frame, offset):
    """/root/reference/src/visualization/overlays.py:103-148 on cv2."""
    h, w = frame.shape[:2]
    indicator_w, indicator_h = 200, 30
    x_start, y_start = (w - indicator_w) // 2, h - 50
    cv2.rectangle(frame, (x_start, y_start), (x_start + indicator_w, y_start + indicator_h), (50, 50, 50), -1)
    cv2.rectangle(frame, (x_start, y_start), (x_start + indicator_w, y_start + indicator_h), (100, 100, 100), 1)
    center_x = x_start + indicator_w // 2
    cv2.line(frame, (center_x, y_start), (center_x, y_start + indicator_h), (255, 255, 255), 1)
    if offset is not None:
        offset_px = int(np.clip(offset, -100, 100))
        color = (0, 255, 0) if abs(offset) < 20 else (0, 255, 255) if abs(offset) < 50 else (0, 0, 255)
        cv2.circle(frame, (center_x + offset_px, y_start + indicator_h // 2), 8, color, -1)
        cv2.putText(frame, f"Offset: {offset:.0f}px", (x_start + 5, y_start - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.4,
                    (255, 255, 255), 1)
    return frame


def lanes_of(points, valid):
    return [points[s] if valid[s] else None for s in range(2)]


def random_mix(rng, recorder_factory, n_cmds=4, max_wh=(90, 70), margin=40):
    """A random image plus a random mix of cv2 drawing calls applied with cv2 (``ref``) and recorded through
    ``recorder_factory() -> DrawList-like`` for frame 0.  Returns (start image, ref image, recorder)."""
    w, h = int(rng.integers(1, max_wh[0])), int(rng.integers(1, max_wh[1]))

    def rp():
        return (int(rng.integers(-margin, w + margin)), int(rng.integers(-margin, h + margin)))

    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = img.copy()
    dl = recorder_factory()
    for _ in range(int(rng.integers(1, n_cmds + 1))):
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        kind = int(rng.integers(0, 7))
        if kind == 0:
            a, b, th = rp(), rp(), int(rng.integers(1, 6))
            cv2.line(ref, a, b, col, th)
            dl.line(0, a, b, col, th)
        elif kind == 1:
            a, b, th = rp(), rp(), int(rng.choice([-1, 1, 2, 3]))
            cv2.rectangle(ref, a, b, col, th)
            dl.rectangle(0, a, b, col, th)
        elif kind == 2:
            c, r = rp(), int(rng.integers(0, 25))
            cv2.circle(ref, c, r, col, -1)
            dl.circle(0, c, r, col, -1)
        elif kind == 3:
            pts = np.array([rp() for _ in range(int(rng.integers(1, 9)))], np.int32)
            cv2.fillPoly(ref, [pts], col)
            dl.fillPoly(0, pts, col)
        elif kind == 4:
            pts = np.array([rp() for _ in range(int(rng.integers(1, 9)))], np.int32)
            th, closed = int(rng.integers(1, 5)), bool(rng.integers(0, 2))
            cv2.polylines(ref, [pts], closed, col, th)
            dl.polylines(0, pts, closed, col, th)
        elif kind == 5:
            pts = np.array([rp() for _ in range(int(rng.integers(1, 9)))], np.int32)
            al, be = [(0.7, 0.3), (0.5, 0.5), (0.3, 0.6)][int(rng.integers(0, 3))]
            o = ref.copy()
            cv2.fillPoly(o, [pts], col)
            ref = cv2.addWeighted(ref, al, o, be, 0)
            dl.fillPoly_weighted(0, pts, col, al, be, 0.0)
        else:
            y0, cnt = int(rng.integers(-5, h)), int(rng.integers(1, 10))
            x1, x2 = int(rng.integers(-10, w + 10)), int(rng.integers(-10, w + 10))
            cols = [tuple(int(c) for c in rng.integers(0, 256, 3)) for _ in range(cnt)]
            for i, c in enumerate(cols):
                cv2.line(ref, (x1, y0 + i), (x2, y0 + i), c, 1)
            dl.rows(0, y0, x1, x2, cols)
    return img, ref, dl


# ---- the CPU replay of K7's device primitives (tests/native/draw_emulator.cpp), built on demand with g++
_EMU = None


def emulator():
    global _EMU
    if _EMU is None:
        import oracle
        _EMU = C.CDLL(oracle.build_draw_emulator())
    return _EMU


def emu_commands(dl, frames):
    """Run a DrawList through the product's host expansion and the CPU replay of the primitives, in place."""
    words, begin = dl.pack()
    n, h, w = frames.shape[:3]
    n_prims = C.c_int64()
    rc = emulator().emu_draw_commands(frames.ctypes.data_as(C.c_void_p), n, h, w, words.ctypes.data_as(C.c_void_p),
                                      begin.ctypes.data_as(C.c_void_p), C.byref(n_prims))
    assert rc == 0
    return n_prims.value


def emu_draw_lanes(frames, points, valid, fill):
    """points int32 [n,2,50,2], valid uint8 [n,2]."""
    n, h, w = frames.shape[:3]
    lp, rp = np.ascontiguousarray(points[:, 0]), np.ascontiguousarray(points[:, 1])
    lv, rv = np.ascontiguousarray(valid[:, 0]), np.ascontiguousarray(valid[:, 1])
    rc = emulator().emu_draw_lanes(frames.ctypes.data_as(C.c_void_p), n, h, w, lp.ctypes.data_as(C.c_void_p),
                                   lv.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p),
                                   rv.ctypes.data_as(C.c_void_p), int(fill))
    assert rc == 0


__all__ = ["draw_golden", "cv2_draw_lanes", "cv2_offset_indicator", "lanes_of", "random_mix", "emulator", "emu_commands",
           "emu_draw_lanes", "h16"]
