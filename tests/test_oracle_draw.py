"""Pins oracle/draw.py (the restatement of OpenCV's rasterisers behind draw_lanes, the offset indicator and the
generator) against cv2 itself -- every primitive on random and degenerate inputs with vertices outside the image -- and
the composed calls against goldens produced by the unmodified reference (tests/golden/make_golden_draw.py)."""
import cv2
import numpy as np
import pytest

from draw_cases import N_RANDOM, STREAMS, random_case
from draw_util import cv2_draw_lanes, cv2_offset_indicator, draw_golden, h16, lanes_of
from oracle import draw as D
from util import gen_frames


def _rp(rng, w, h, m=40):
    return (int(rng.integers(-m, w + m)), int(rng.integers(-m, h + m)))


def _size(rng):
    return int(rng.integers(1, 80)), int(rng.integers(1, 60))


def test_thin_and_thick_lines_match_cv2():
    rng = np.random.default_rng(0)
    for th in (1, 2, 3, 4, 5):
        for _ in range(400):
            w, h = _size(rng)
            a, b = _rp(rng, w, h), _rp(rng, w, h)
            i1 = np.zeros((h, w, 3), np.uint8)
            i2 = i1.copy()
            cv2.line(i1, a, b, (1, 2, 3), th)
            D.line(i2, a, b, (1, 2, 3), th)
            assert np.array_equal(i1, i2), (th, w, h, a, b)


def test_line_closed_form_equals_the_error_recurrence():
    # the device kernel places point i of a Bresenham line with floor((2*dmin*i + dmaj - 1) / (2*dmaj))
    rng = np.random.default_rng(1)
    for _ in range(300):
        w, h = int(rng.integers(2, 300)), int(rng.integers(2, 300))
        p1, p2 = _rp(rng, w, h, 0), _rp(rng, w, h, 0)
        p1 = (min(max(p1[0], 0), w - 1), min(max(p1[1], 0), h - 1))
        p2 = (min(max(p2[0], 0), w - 1), min(max(p2[1], 0), h - 1))
        pts = D.line_points(w, h, p1, p2)
        x1, y1 = pts[0]
        dx, dy = abs(pts[-1][0] - x1), abs(pts[-1][1] - y1)
        sy = 1 if pts[-1][1] >= y1 else -1
        vert = dy > dx
        dmaj, dmin = (dy, dx) if vert else (dx, dy)
        for i, (x, y) in enumerate(pts):
            k = (2 * dmin * i + dmaj - 1) // (2 * dmaj) if dmaj else 0
            assert (x, y) == ((x1 + k, y1 + sy * i) if vert else (x1 + i, y1 + sy * k))


def test_polygon_outline_dda_matches_two_vertex_fillconvexpoly():
    # cv2.fillConvexPoly with two 16.16 vertices draws only its outline: Line2 there and back
    rng = np.random.default_rng(2)
    for _ in range(1500):
        w, h = _size(rng)
        pts = [(int(rng.integers(-30 * 65536, (w + 30) * 65536)), int(rng.integers(-30 * 65536, (h + 30) * 65536)))
               for _ in range(2)]
        i1 = np.zeros((h, w), np.uint8)
        i2 = i1.copy()
        cv2.fillConvexPoly(i1, np.array(pts, np.int32), 1, 8, 16)
        D.line2(i2, pts[1], pts[0], 1)
        D.line2(i2, pts[0], pts[1], 1)
        assert np.array_equal(i1, i2), (w, h, pts)


def test_thin_line_with_fractional_end_points_is_bresenham_on_rounded_points():
    rng = np.random.default_rng(3)
    for _ in range(500):
        w, h = _size(rng)
        pts = [(int(rng.integers(-30 * 65536, (w + 30) * 65536)), int(rng.integers(-30 * 65536, (h + 30) * 65536)))
               for _ in range(2)]
        i1 = np.zeros((h, w), np.uint8)
        i2 = i1.copy()
        cv2.line(i1, pts[0], pts[1], 1, 1, 8, 16)
        D.thin_line_shifted(i2, pts[0], pts[1], 1)
        assert np.array_equal(i1, i2)


@pytest.mark.parametrize("shift", [0, 16])
def test_fill_convex_poly_matches_cv2(shift):
    rng = np.random.default_rng(4 + shift)
    for _ in range(1000):
        w, h = _size(rng)
        s = 1 << shift
        pts = np.array([(int(rng.integers(-30 * s, (w + 30) * s)), int(rng.integers(-30 * s, (h + 30) * s)))
                        for _ in range(3)], np.int32)
        i1 = np.zeros((h, w), np.uint8)
        i2 = i1.copy()
        cv2.fillConvexPoly(i1, pts, 1, 8, shift)
        D.fill_convex_poly(i2, pts, 1, shift)
        assert np.array_equal(i1, i2), (w, h, pts.tolist())


def test_circles_and_rectangles_match_cv2():
    rng = np.random.default_rng(5)
    for _ in range(800):
        w, h = _size(rng)
        c, r = _rp(rng, w, h), int(rng.integers(0, 30))
        i1 = np.zeros((h, w, 3), np.uint8)
        i2 = i1.copy()
        cv2.circle(i1, c, r, (1, 2, 3), -1)
        D.circle(i2, c, r, (1, 2, 3), -1)
        assert np.array_equal(i1, i2), (w, h, c, r)
        a, b = _rp(rng, w, h), _rp(rng, w, h)
        for th in (-1, 1, 2, 7):
            i1 = np.zeros((h, w, 3), np.uint8)
            i2 = i1.copy()
            cv2.rectangle(i1, a, b, (1, 2, 3), th)
            D.rectangle(i2, a, b, (1, 2, 3), th)
            assert np.array_equal(i1, i2), (w, h, a, b, th)


def test_fillpoly_matches_cv2_including_self_intersections_and_outside_vertices():
    rng = np.random.default_rng(6)
    for t in range(2500):
        w, h = _size(rng)
        n = int(rng.integers(1, 9))
        m = 40 if t % 3 else 0
        pts = np.array([_rp(rng, w, h, m) for _ in range(n)], np.int32)
        i1 = np.zeros((h, w, 3), np.uint8)
        i2 = i1.copy()
        cv2.fillPoly(i1, [pts], (1, 2, 3))
        D.fill_poly(i2, pts, (1, 2, 3))
        assert np.array_equal(i1, i2), (w, h, pts.tolist())


def test_polylines_match_cv2():
    rng = np.random.default_rng(7)
    for _ in range(700):
        w, h = _size(rng)
        pts = np.array([_rp(rng, w, h) for _ in range(int(rng.integers(1, 9)))], np.int32)
        for th in (1, 2, 3):
            for closed in (False, True):
                i1 = np.zeros((h, w), np.uint8)
                i2 = i1.copy()
                cv2.polylines(i1, [pts], closed, 1, th)
                D.polylines(i2, pts, closed, 1, th)
                assert np.array_equal(i1, i2), (th, closed, w, h, pts.tolist())


def test_add_weighted_is_two_float32_fmas_on_every_value_pair():
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    b = a.T.copy()
    for al, be, ga in [(0.7, 0.3, 0.0), (0.5, 0.5, 0.0), (0.3, 0.6, 0.0), (0.45, 0.8, 0.0)]:
        assert np.array_equal(cv2.addWeighted(a, al, b, be, ga), D.add_weighted_u8(a, al, b, be, ga)), (al, be, ga)
    # pixels outside the lane polygon keep their value: addWeighted(v, 0.7, v, 0.3, 0) == v
    v = np.arange(256, dtype=np.uint8)
    assert np.array_equal(D.add_weighted_u8(v, 0.7, v, 0.3), v)
    # no scalar-tail effect for gamma == 0: odd lengths, random data
    rng = np.random.default_rng(8)
    for total in (1, 7, 63, 1001, 1063):
        x = rng.integers(0, 256, (40, total), dtype=np.uint8)
        y = rng.integers(0, 256, (40, total), dtype=np.uint8)
        assert np.array_equal(cv2.addWeighted(x, 0.7, y, 0.3, 0), D.add_weighted_u8(x, 0.7, y, 0.3))


def _text_mask(text, org, shape):
    canvas = np.zeros(shape[:2], np.uint8)
    cv2.putText(canvas, text, org, cv2.FONT_HERSHEY_SIMPLEX, 0.4, 255, 1)
    return canvas > 0


def test_composed_calls_match_the_reference_goldens():
    g = draw_golden()
    for w, h, n in STREAMS:
        key = f"stream_{w}x{h}"
        frames = gen_frames(w, h, n)
        for i in range(n):
            if (w, i) not in ((640, 0), (640, 5), (640, 11), (1920, 1)):      # pure-Python rasteriser: a few frames
                continue
            l, r = lanes_of(g[key + "_points"][i], g[key + "_valid"][i])
            off = None if np.isnan(g[key + "_offset"][i]) else float(g[key + "_offset"][i])
            filled = D.draw_lanes(frames[i], l, r, True)
            assert h16(filled) == g[key + "_hash"][i][0]
            assert h16(D.draw_lanes(frames[i], l, r, False)) == g[key + "_hash"][i][1]
            assert h16(D.draw_lane_offset_indicator(filled.copy(), off, _text_mask)) == g[key + "_hash"][i][2]
    for seed in range(N_RANDOM):
        frame, pts, valid, off = random_case(seed)
        l, r = lanes_of(pts, valid)
        filled = D.draw_lanes(frame, l, r, True)
        assert h16(filled) == g["random_hash"][seed][0], seed
        assert h16(D.draw_lanes(frame, l, r, False)) == g["random_hash"][seed][1], seed
        assert h16(D.draw_lane_offset_indicator(filled.copy(), off, _text_mask)) == g["random_hash"][seed][2], seed


def test_the_cv2_call_sequences_of_the_tests_equal_the_reference_goldens():
    # draw_util.cv2_draw_lanes / cv2_offset_indicator are what the GPU tests compare with where no golden exists
    g = draw_golden()
    for seed in range(N_RANDOM):
        frame, pts, valid, off = random_case(seed)
        l, r = lanes_of(pts, valid)
        filled = cv2_draw_lanes(frame.copy(), l, r, True)
        assert h16(filled) == g["random_hash"][seed][0]
        assert h16(cv2_draw_lanes(frame.copy(), l, r, False)) == g["random_hash"][seed][1]
        assert h16(cv2_offset_indicator(filled.copy(), off)) == g["random_hash"][seed][2]
