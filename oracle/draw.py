"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the OpenCV rasterisers behind the reference's drawing calls.

SURVEY.md section 8(f) ranks 3 and 4: the step right after the lane path is ``LaneDetector.draw_lanes``
(/root/reference/src/perception/lane_detector.py:220-251: ``cv2.fillPoly`` on a copy, ``cv2.addWeighted``,
``cv2.polylines`` thickness 3) and ``OverlayRenderer.draw_lane_offset_indicator``
(/root/reference/src/visualization/overlays.py:103-148: filled / outlined ``cv2.rectangle``, ``cv2.line``, filled
``cv2.circle``, ``cv2.putText``); the input fixture of every config is the reference's ``SyntheticDataGenerator``
(bytecode only), whose frames are made of ``cv2.line`` (thickness 1 and 2), ``cv2.rectangle``,
``cv2.fillPoly`` and filled ``cv2.circle``.

The arithmetic lives in OpenCV (opencv-python >= 4.5.0, installed 4.13.0.92; source not on this box).  This file
restates the published algorithms of ``modules/imgproc/src/drawing.cpp`` for 8-bit images and ``LINE_8``:
``clipLine``, ``LineIterator`` (8-connected Bresenham, left-to-right), ``Line2`` (thin line with 16.16 fixed-point end points),
``FillConvexPoly``, ``Circle``, ``ThickLine``, ``PolyLine``, ``CollectPolyEdges`` + ``FillEdgeCollection``
(``fillPoly``), ``rectangle`` and ``addWeighted`` on uint8.  tests/test_oracle_draw.py pins every function against
cv2 itself on random and degenerate inputs (vertices outside the image included), and the composed ``draw_lanes`` /
``draw_lane_offset_indicator`` / generator frames against the reference's own code.

``putText`` is not restated (the Hershey glyph tables are data of the OpenCV build): text enters as a bit mask
rendered once per distinct string by ``cv2.putText`` itself, the same way the ROI polygon mask enters the lane path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT
DBL_EPSILON = 2.220446049250313e-16


def _trunc(v: float) -> int:
    """C's (int64)(double) conversion."""
    return int(v)


def _cv_round(v: float) -> int:
    """cvRound on a double: round half to even (lrint in the default rounding mode)."""
    return int(np.rint(v))


# ------------------------------------------------------------------------------------------------ clipLine
def clip_line(width: int, height: int, p1, p2):
    """``cv::clipLine(Size2l, Point2l&, Point2l&)``: returns (visible, p1, p2)."""
    x1, y1 = int(p1[0]), int(p1[1])
    x2, y2 = int(p2[0]), int(p2[1])
    right, bottom = width - 1, height - 1
    if width <= 0 or height <= 0:
        return False, (x1, y1), (x2, y2)
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _trunc(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _trunc(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _trunc(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _trunc(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


# ------------------------------------------------------------------------------------------------ LineIterator / Line
def line_points(width: int, height: int, p1, p2):
    """Pixels ``Line(img, p1, p2, color, 8)`` writes: 8-connected Bresenham between the clipped end points, walked
    left to right.  Closed form per point: the minor coordinate has advanced ``floor((2*dmin*i + dmaj - 1) / (2*dmaj))``
    times after ``i`` major steps (what the device kernel evaluates); this routine runs the error recurrence."""
    x1, y1 = int(p1[0]), int(p1[1])
    x2, y2 = int(p2[0]), int(p2[1])
    if not (0 <= x1 < width and 0 <= x2 < width and 0 <= y1 < height and 0 <= y2 < height):
        ok, (x1, y1), (x2, y2) = clip_line(width, height, (x1, y1), (x2, y2))
        if not ok:
            return []
    dx, dy = x2 - x1, y2 - y1
    sx = sy = 1
    if dx < 0:
        dx, dy = -dx, -dy
        x1, y1 = x2, y2
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
    err = dx - 2 * dy
    pts = []
    x, y = x1, y1
    for _ in range(dx + 1):
        pts.append((x, y))
        step = err < 0
        err += -2 * dy + (2 * dx if step else 0)
        if vert:
            y += sy
            x += sx if step else 0
        else:
            x += sx
            y += sy if step else 0
    return pts


def line8(img, p1, p2, color):
    h, w = img.shape[:2]
    for x, y in line_points(w, h, p1, p2):
        img[y, x] = color


# ------------------------------------------------------------------------------------------------ Line2
def line2(img, p1, p2, color):
    """``Line2``: end points in 16.16 fixed point, a DDA along the major axis.  This is what ``FillConvexPoly`` outlines
    its polygon with when ``shift != 0`` (every thick line); pinned against two-vertex ``cv2.fillConvexPoly(shift=16)``
    calls, which draw only that outline."""
    h, w = img.shape[:2]
    ok, (x1, y1), (x2, y2) = clip_line(w << XY_SHIFT, h << XY_SHIFT, p1, p2)
    if not ok:
        return
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)

    def put(x, y):
        if 0 <= x < w and 0 <= y < h:
            img[y, x] = color

    if ax > ay:
        if dx < 0:
            dy = -dy
            x1, x2 = x2, x1
            y1, y2 = y2, y1
        y_step = _cdiv(dy * XY_ONE, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        if dy < 0:
            dx = -dx
            x1, x2 = x2, x1
            y1, y2 = y2, y1
        x_step = _cdiv(dx * XY_ONE, ay | 1)
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    put((x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT)
    if ax > ay:
        x1 >>= XY_SHIFT
        for i in range(ecount + 1):
            put(x1 + i, (y1 + i * y_step) >> XY_SHIFT)
    else:
        y1 >>= XY_SHIFT
        for i in range(ecount + 1):
            put((x1 + i * x_step) >> XY_SHIFT, y1 + i)


def thin_line_shifted(img, p1, p2, color):
    """``cv2.line(img, p1, p2, color, 1, LINE_8, shift=16)``: OpenCV 4.13 draws it as ``Line`` between the end points
    rounded to whole pixels (pinned on 4000 random lines).  Not used by the reference; kept because it documents that
    the thin-line path and the polygon outline are different rasterisers."""
    half = XY_ONE >> 1
    line8(img, ((int(p1[0]) + half) >> XY_SHIFT, (int(p1[1]) + half) >> XY_SHIFT),
          ((int(p2[0]) + half) >> XY_SHIFT, (int(p2[1]) + half) >> XY_SHIFT), color)


def _cdiv(a: int, b: int) -> int:
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def _hline(img, y, x1, x2, color):
    if x2 >= x1:
        img[y, x1:x2 + 1] = color


# ------------------------------------------------------------------------------------------------ FillConvexPoly
def fill_convex_poly(img, v, color, shift=0):
    """``FillConvexPoly(img, v, npts, color, LINE_8, shift)``; ``v`` in ``shift``-bit fixed point."""
    h, w = img.shape[:2]
    v = [(int(x), int(y)) for x, y in v]
    npts = len(v)
    delta = (1 << shift) >> 1
    delta1 = delta2 = XY_ONE >> 1
    p0 = (v[-1][0] << (XY_SHIFT - shift), v[-1][1] << (XY_SHIFT - shift))
    xmin = xmax = v[0][0]
    ymin = ymax = v[0][1]
    imin = 0
    for i, (px, py) in enumerate(v):
        if py < ymin:
            ymin, imin = py, i
        ymax = max(ymax, py)
        xmax = max(xmax, px)
        xmin = min(xmin, px)
        p = (px << (XY_SHIFT - shift), py << (XY_SHIFT - shift))
        if shift == 0:
            line8(img, (p0[0] >> XY_SHIFT, p0[1] >> XY_SHIFT), (p[0] >> XY_SHIFT, p[1] >> XY_SHIFT), color)
        else:
            line2(img, p0, p, color)
        p0 = p
    xmin = (xmin + delta) >> shift
    xmax = (xmax + delta) >> shift
    ymin = (ymin + delta) >> shift
    ymax = (ymax + delta) >> shift
    if npts < 3 or xmax < 0 or ymax < 0 or xmin >= w or ymin >= h:
        return
    ymax = min(ymax, h - 1)
    edge = [dict(idx=imin, di=1, x=-XY_ONE, dx=0, ye=ymin), dict(idx=imin, di=npts - 1, x=-XY_ONE, dx=0, ye=ymin)]
    edges = npts
    y = ymin
    while True:
        for e in edge:
            if y >= e["ye"]:
                idx0, di = e["idx"], e["di"]
                idx = idx0 + di
                if idx >= npts:
                    idx -= npts
                while True:
                    edges -= 1
                    if edges < 0:       # `for (; edges-- > 0; )` leaves edges at -1 on exhaustion
                        break
                    ty = (v[idx][1] + delta) >> shift
                    if ty > y:
                        xs = v[idx0][0] << (XY_SHIFT - shift)
                        xe = v[idx][0] << (XY_SHIFT - shift)
                        e["ye"] = ty
                        e["dx"] = _cdiv((xe - xs) * 2 + (ty - y), 2 * (ty - y))
                        e["x"] = xs
                        e["idx"] = idx
                        break
                    idx0 = idx
                    idx += di
                    if idx >= npts:
                        idx -= npts
        if edges < 0:
            break
        if y >= 0:
            l, r = (1, 0) if edge[0]["x"] > edge[1]["x"] else (0, 1)
            xx1 = (edge[l]["x"] + delta1) >> XY_SHIFT
            xx2 = (edge[r]["x"] + delta2) >> XY_SHIFT
            if xx2 >= 0 and xx1 < w:
                _hline(img, y, max(xx1, 0), min(xx2, w - 1), color)
        edge[0]["x"] += edge[0]["dx"]
        edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break


# ------------------------------------------------------------------------------------------------ Circle
def circle_spans(radius: int):
    """Row spans of ``Circle(img, c, radius, color, fill=1)`` relative to the centre: list of (dy, dx) meaning the row
    ``cy + dy`` is filled over ``[cx - dx, cx + dx]`` (rows repeat; the union is what is drawn)."""
    spans = []
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        spans += [(-dy, dx), (dy, dx), (-dx, dy), (dx, dy)]
        dy += 1
        err += plus
        plus += 2
        mask = -1 if err > 0 else 0
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return spans


def circle_filled(img, center, radius, color):
    h, w = img.shape[:2]
    cx, cy = int(center[0]), int(center[1])
    for dy, dx in circle_spans(radius):
        y, x1, x2 = cy + dy, cx - dx, cx + dx
        if 0 <= y < h and x1 < w and x2 >= 0:
            _hline(img, y, max(x1, 0), min(x2, w - 1), color)


# ------------------------------------------------------------------------------------------------ ThickLine / PolyLine
def thick_line(img, p0, p1, color, thickness, flags=3, shift=0):
    """``ThickLine(img, p0, p1, color, thickness, LINE_8, flags, shift)``.  OpenCV 4.13 first clips a thick segment
    (thickness > 1) against the image rectangle grown by ``thickness`` pixels on every side, with ``clipLine``'s
    truncating arithmetic, and drops it when nothing is left -- found by probing and pinned on random lines, polylines
    and rectangles with vertices outside the image (tests/test_oracle_draw.py); it changes the slope of a clipped
    segment's quadrilateral by a few 2^-16."""
    if thickness > 1 and shift == 0:
        h, w = img.shape[:2]
        m = thickness
        ok, a, b = clip_line(w + 2 * m, h + 2 * m, (int(p0[0]) + m, int(p0[1]) + m), (int(p1[0]) + m, int(p1[1]) + m))
        if not ok:
            return
        p0, p1 = (a[0] - m, a[1] - m), (b[0] - m, b[1] - m)
    p0 = (int(p0[0]) << (XY_SHIFT - shift), int(p0[1]) << (XY_SHIFT - shift))
    p1 = (int(p1[0]) << (XY_SHIFT - shift), int(p1[1]) << (XY_SHIFT - shift))
    if thickness <= 1:
        if shift == 0:
            line8(img, ((p0[0] + (XY_ONE >> 1)) >> XY_SHIFT, (p0[1] + (XY_ONE >> 1)) >> XY_SHIFT),
                  ((p1[0] + (XY_ONE >> 1)) >> XY_SHIFT, (p1[1] + (XY_ONE >> 1)) >> XY_SHIFT), color)
        else:
            thin_line_shifted(img, p0, p1, color)
        return
    inv = 1.0 / XY_ONE
    dx = (p0[0] - p1[0]) * inv
    dy = (p1[1] - p0[1]) * inv
    r = dx * dx + dy * dy
    odd = thickness & 1
    thickness <<= XY_SHIFT - 1
    if abs(r) > DBL_EPSILON:
        r = (thickness + odd * XY_ONE * 0.5) / math.sqrt(r)
        dpx, dpy = _cv_round(dy * r), _cv_round(dx * r)
        quad = [(p0[0] + dpx, p0[1] + dpy), (p0[0] - dpx, p0[1] - dpy), (p1[0] - dpx, p1[1] - dpy),
                (p1[0] + dpx, p1[1] + dpy)]
        fill_convex_poly(img, quad, color, XY_SHIFT)
    for i in range(2):
        if flags & (i + 1):
            c = ((p0[0] + (XY_ONE >> 1)) >> XY_SHIFT, (p0[1] + (XY_ONE >> 1)) >> XY_SHIFT)
            circle_filled(img, c, (thickness + (XY_ONE >> 1)) >> XY_SHIFT, color)
        p0 = p1


def polylines(img, pts, closed, color, thickness=1):
    """``cv2.polylines(img, [pts], closed, color, thickness)``."""
    pts = [(int(x), int(y)) for x, y in np.asarray(pts).reshape(-1, 2)]
    if not pts:
        return
    count = len(pts)
    i = count - 1 if closed else 0
    flags = 2 + (0 if closed else 1)
    p0 = pts[i]
    for i in range(0 if closed else 1, count):
        thick_line(img, p0, pts[i], color, thickness, flags)
        p0 = pts[i]
        flags = 2


def line(img, p1, p2, color, thickness=1):
    """``cv2.line(img, p1, p2, color, thickness)``."""
    thick_line(img, p1, p2, color, thickness, 3)


def rectangle(img, p1, p2, color, thickness=1):
    """``cv2.rectangle(img, p1, p2, color, thickness)`` (thickness < 0: filled)."""
    pt = [(p1[0], p1[1]), (p2[0], p1[1]), (p2[0], p2[1]), (p1[0], p2[1])]
    if thickness >= 0:
        polylines(img, pt, True, color, thickness)
    else:
        fill_convex_poly(img, pt, color, 0)


def circle(img, center, radius, color, thickness=-1):
    """``cv2.circle`` for the filled case the reference uses."""
    if thickness >= 0:
        raise NotImplementedError("only filled circles are on the path")
    circle_filled(img, center, radius, color)


# ------------------------------------------------------------------------------------------------ fillPoly
def poly_edges(width: int, height: int, pts):
    """``CollectPolyEdges`` for LINE_8, shift 0, no offset: returns (boundary segments ``Line`` draws, edge table).
    Edge = dict(y0, y1, x, dx): x in 16.16 at row y0, dx per row, active on rows y0 <= y < y1."""
    v = [(int(x), int(y)) for x, y in np.asarray(pts).reshape(-1, 2)]
    segs, edges = [], []
    count = len(v)
    pt0 = (v[-1][0] << XY_SHIFT, v[-1][1])
    for i in range(count):
        pt1 = (v[i][0] << XY_SHIFT, v[i][1])
        t0 = ((pt0[0] + (XY_ONE >> 1)) >> XY_SHIFT, pt0[1])
        t1 = ((pt1[0] + (XY_ONE >> 1)) >> XY_SHIFT, pt1[1])
        segs.append((t0, t1))
        pt0c, pt1c = list(pt0), list(pt1)
        if not (0 <= t0[0] < width and 0 <= t1[0] < width and 0 <= t0[1] < height and 0 <= t1[1] < height):
            _, c0, c1 = clip_line(width, height, t0, t1)
            pt0c[0], pt1c[0] = c0[0] << XY_SHIFT, c1[0] << XY_SHIFT
            if c0[1] != c1[1]:
                pt0c[1], pt1c[1] = c0[1], c1[1]
        if pt0[1] != pt1[1]:
            dx = _cdiv(pt1c[0] - pt0c[0], pt1c[1] - pt0c[1])
            if pt0[1] < pt1[1]:
                e = dict(y0=pt0[1], y1=pt1[1], x=pt0c[0] + (pt0[1] - pt0c[1]) * dx, dx=dx)
            else:
                e = dict(y0=pt1[1], y1=pt0[1], x=pt1c[0] + (pt1[1] - pt1c[1]) * dx, dx=dx)
            edges.append(e)
        pt0 = pt1
    return segs, edges


def fill_edge_rows(width: int, height: int, edges):
    """``FillEdgeCollection`` (LINE_8) as a per-row rule: on row y the active edges (y0 <= y < y1) are ordered by their
    x = x0 + (y - y0)*dx and paired (0,1), (2,3), ...; a pair fills [xa >> 16, xb >> 16].  Returns {y: [(x1, x2), ...]}
    of clipped spans.  (The scan-line code keeps an active list sorted by a bubble sort at the end of every row and
    inserts new edges in front of the first entry that is not smaller, which is exactly this order; a closed polygon
    crosses every row an even number of times.)"""
    rows = {}
    if len(edges) < 2:
        return rows
    y_min = min(e["y0"] for e in edges)
    y_max = max(e["y1"] for e in edges)
    xs = [e["x"] for e in edges] + [e["x"] + (e["y1"] - e["y0"]) * e["dx"] for e in edges]
    if y_max < 0 or y_min >= height or max(xs) < 0 or min(xs) >= (width << XY_SHIFT):
        return rows
    for y in range(max(y_min, 0), min(y_max, height)):
        act = sorted(e["x"] + (y - e["y0"]) * e["dx"] for e in edges if e["y0"] <= y < e["y1"])
        spans = []
        for k in range(0, len(act) - 1, 2):
            x1, x2 = (act[k] + XY_ONE - 1) >> XY_SHIFT, act[k + 1] >> XY_SHIFT
            if x1 < width and x2 >= 0:
                spans.append((max(x1, 0), min(x2, width - 1)))
        if spans:
            rows[y] = spans
    return rows


def fill_poly(img, pts, color):
    """``cv2.fillPoly(img, [pts], color)`` (one contour, LINE_8)."""
    h, w = img.shape[:2]
    segs, edges = poly_edges(w, h, pts)
    for t0, t1 in segs:
        line8(img, t0, t1, color)
    for y, spans in fill_edge_rows(w, h, edges).items():
        for x1, x2 in spans:
            _hline(img, y, x1, x2, color)


def fill_poly_mask(width: int, height: int, pts) -> np.ndarray:
    m = np.zeros((height, width), np.uint8)
    fill_poly(m, pts, 1)
    return m.astype(bool)


# ------------------------------------------------------------------------------------------------ addWeighted
def add_weighted_u8(a: np.ndarray, alpha: float, b: np.ndarray, beta: float, gamma: float = 0.0) -> np.ndarray:
    """``cv2.addWeighted`` on uint8 as this build computes it: float32 ``fma(a, alpha, fma(b, beta, gamma))`` (two
    fused multiply-adds, scalars converted to float32 first), rounded half to even and saturated.  Pinned on all
    65 536 (a, b) pairs for three scalar sets; the unfused float32 form differs on 128 pairs for (0.7, 0.3, 0).
    The fused operations are evaluated in float64 here: a product of an 8-bit and a 24-bit significand plus a float32
    addend is exact in 53 bits, so the final conversion to float32 is the single rounding of the fma."""
    fa, fb, fg = float(np.float32(alpha)), float(np.float32(beta)), float(np.float32(gamma))
    inner = (b.astype(np.float64) * fb + fg).astype(np.float32)
    t = (a.astype(np.float64) * fa + inner.astype(np.float64)).astype(np.float32)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ composed calls
LANE_FILL_COLOR = (0, 255, 100)        # lane_detector.py:243
LEFT_COLOR = (255, 0, 0)               # :248
RIGHT_COLOR = (0, 0, 255)              # :251


def draw_lanes(frame, left_points, right_points, fill_lane=True):
    """``LaneDetector.draw_lanes`` (lane_detector.py:220-251) on the restated rasterisers.  ``*_points``: int32[50,2] or
    None.  Returns a new frame (the reference draws in place when the fill branch is skipped; callers compare values)."""
    frame = frame.copy()
    if fill_lane and left_points is not None and right_points is not None:
        overlay = frame.copy()
        pts = np.vstack([left_points, right_points[::-1]])
        fill_poly(overlay, pts, LANE_FILL_COLOR)
        frame = add_weighted_u8(frame, 0.7, overlay, 0.3, 0.0)
    if left_points is not None:
        polylines(frame, left_points, False, LEFT_COLOR, 3)
    if right_points is not None:
        polylines(frame, right_points, False, RIGHT_COLOR, 3)
    return frame


def offset_indicator_geometry(width: int, height: int, offset):
    """Scalars of ``draw_lane_offset_indicator`` (overlays.py:103-148)."""
    iw, ih = 200, 30
    x0 = (width - iw) // 2
    y0 = height - 50
    cx = x0 + iw // 2
    g = dict(x0=x0, y0=y0, x1=x0 + iw, y1=y0 + ih, cx=cx, dot=None, color=None, text=None, text_org=(x0 + 5, y0 - 5))
    if offset is not None:
        off_px = int(np.clip(offset, -100, 100))
        g["dot"] = (cx + off_px, y0 + ih // 2)
        a = abs(offset)
        g["color"] = (0, 255, 0) if a < 20 else (0, 255, 255) if a < 50 else (0, 0, 255)
        g["text"] = f"Offset: {offset:.0f}px"
    return g


def draw_lane_offset_indicator(frame, offset, text_mask_fn):
    """``OverlayRenderer.draw_lane_offset_indicator`` in place; ``text_mask_fn(text, org, shape) -> bool[H,W]`` supplies
    the pixels ``cv2.putText(..., FONT_HERSHEY_SIMPLEX, 0.4, ..., 1)`` sets."""
    h, w = frame.shape[:2]
    g = offset_indicator_geometry(w, h, offset)
    rectangle(frame, (g["x0"], g["y0"]), (g["x1"], g["y1"]), (50, 50, 50), -1)
    rectangle(frame, (g["x0"], g["y0"]), (g["x1"], g["y1"]), (100, 100, 100), 1)
    line(frame, (g["cx"], g["y0"]), (g["cx"], g["y1"]), (255, 255, 255), 1)
    if offset is not None:
        circle_filled(frame, g["dot"], 8, g["color"])
        frame[text_mask_fn(g["text"], g["text_org"], frame.shape)] = (255, 255, 255)
    return frame
