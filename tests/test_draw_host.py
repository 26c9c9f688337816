"""Host half of the device rasteriser (csrc/draw_prims.h: cv2-level calls -> device primitives) without a GPU: the
primitives are replayed on the CPU by tests/native/draw_emulator.cpp with the kernel's per-primitive rules and the
pixels are compared with cv2 itself, with the reference goldens and with the host generator."""
import numpy as np
import pytest

from draw_cases import N_RANDOM, STREAMS, random_case
from draw_util import (cv2_draw_lanes, cv2_offset_indicator, draw_golden, emu_commands, emu_draw_lanes, h16, random_mix)
from multimodal_autonomous_driving_perception_and_planning_b200.generators.synthetic_data import (SyntheticDataGenerator,
                                                                                                  _ListCanvas)
from multimodal_autonomous_driving_perception_and_planning_b200.visualization import DrawList, OverlayRenderer
from util import gen_frames


def test_random_command_mixes_equal_cv2():
    rng = np.random.default_rng(0)
    for t in range(2500):
        img, ref, dl = random_mix(rng, lambda: DrawList(1))
        mine = img.copy()[None]
        emu_commands(dl, mine)
        assert np.array_equal(ref, mine[0]), t


def test_generator_frames_recorded_and_replayed_equal_the_host_generator():
    for w, h, n, start in [(640, 480, 8, 0), (1920, 1080, 2, 95), (1280, 720, 2, 1000), (321, 203, 3, 7)]:
        ref = SyntheticDataGenerator(w, h).generate_batch(n, start_frame=start)
        gen = SyntheticDataGenerator(w, h)
        gen.frame_count = start
        dl = DrawList(n)
        for i in range(n):
            gen._paint_frame_with_vehicles(_ListCanvas(dl, i))
        mine = np.zeros_like(ref)
        emu_commands(dl, mine)
        assert np.array_equal(ref, mine), (w, h)
        assert gen.frame_count == start + n


def test_generator_scene_in_the_library_equals_the_host_generator():
    """lane_generate_frames' host half (C++ scene code with NumPy's legacy RandomState) on the CPU: every seed
    (frame_count % 100) and the sha256 fixtures of SURVEY.md 8(c)."""
    import ctypes as C
    import hashlib
    from draw_util import emulator
    emu = emulator()
    emu.emu_generate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64]
    for w, h, n, start in [(640, 480, 300, 0), (320, 240, 220, 950), (321, 203, 60, 7), (64, 48, 30, 0)]:
        ref = SyntheticDataGenerator(w, h).generate_batch(n, start_frame=start)
        mine = np.full_like(ref, 9)[:, :, :, :] * 0
        assert emu.emu_generate(mine.ctypes.data_as(C.c_void_p), n, h, w, start) == 0
        assert np.array_equal(ref, mine), (w, h)
        if (w, h, start) == (640, 480, 0):
            assert hashlib.sha256(mine.tobytes()).hexdigest()[:16] == "701c6dd0c4e8d707"


def test_draw_lanes_and_offset_indicator_equal_the_reference_goldens():
    g = draw_golden()
    ov = OverlayRenderer()
    for w, h, n in STREAMS:
        key = f"stream_{w}x{h}"
        frames = np.stack(gen_frames(w, h, n))
        for fill, col in ((True, 0), (False, 1)):
            mine = frames.copy()
            emu_draw_lanes(mine, g[key + "_points"], g[key + "_valid"], fill)
            assert [h16(f) for f in mine] == [r[col] for r in g[key + "_hash"]]
            if fill:
                dl = DrawList(n)
                for i in range(n):
                    off = g[key + "_offset"][i]
                    ov.record_lane_offset_indicator(dl, i, w, h, None if np.isnan(off) else float(off))
                emu_commands(dl, mine)
                assert [h16(f) for f in mine] == [r[2] for r in g[key + "_hash"]]
    for seed in range(N_RANDOM):
        frame, pts, valid, off = random_case(seed)
        h, w = frame.shape[:2]
        for fill, col in ((True, 0), (False, 1)):
            mine = frame.copy()[None]
            emu_draw_lanes(mine, pts[None], valid[None], fill)
            assert h16(mine[0]) == g["random_hash"][seed][col], (seed, fill)
        mine = frame.copy()[None]
        emu_draw_lanes(mine, pts[None], valid[None], True)
        dl = DrawList(1)
        ov.record_lane_offset_indicator(dl, 0, w, h, off)
        emu_commands(dl, mine)
        assert h16(mine[0]) == g["random_hash"][seed][2], seed


def test_lanes_far_outside_and_degenerate_polygons():
    rng = np.random.default_rng(3)
    for t in range(200):
        h, w = int(rng.integers(8, 200)), int(rng.integers(8, 260))
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        pts = np.zeros((2, 50, 2), np.int32)
        y = np.linspace(0.6 * h, h, 50)
        for s in range(2):
            kind = t % 4
            if kind == 0:      # both sides on one vertical line: zero-area polygon
                x = np.full(50, w // 2)
            elif kind == 1:    # far outside on one side
                x = np.full(50, (-3000 if s == 0 else 4000)) + rng.integers(-50, 50, 50)
            elif kind == 2:    # crossing lanes
                x = np.linspace(0, w, 50) if s == 0 else np.linspace(w, 0, 50)
            else:              # wild quadratic
                x = np.polyval([rng.normal(0, 0.05), rng.normal(0, 3), rng.normal(w / 2, w)], y)
            pts[s] = np.column_stack([x, y]).astype(np.int32)
        valid = np.ones(2, np.uint8)
        ref = cv2_draw_lanes(frame.copy(), pts[0], pts[1], True)
        mine = frame.copy()[None]
        emu_draw_lanes(mine, pts[None], valid[None], True)
        assert np.array_equal(ref, mine[0]), t


def test_text_cut_by_the_border_is_rendered_in_place():
    ov = OverlayRenderer()
    rng = np.random.default_rng(4)
    for h, w in [(60, 210), (56, 120), (40, 100), (52, 260)]:
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2_offset_indicator(frame.copy(), -37.5)
        dl = DrawList(1)
        ov.record_lane_offset_indicator(dl, 0, w, h, -37.5)
        mine = frame.copy()[None]
        emu_commands(dl, mine)
        assert np.array_equal(ref, mine[0]), (h, w)


def test_malformed_command_streams_are_rejected():
    from draw_util import emulator
    import ctypes as C
    emu = emulator()
    frames = np.zeros((1, 8, 8, 3), np.uint8)
    for words in ([99, 0, 0], [1, 0, 0, 5, 5], [4, 0, 7, 1, 1], [6, 0, 0, 0, 0x40800000, 0], [3, 1, 1, 2, 0, 1]):
        w = np.array(words, np.int32)
        begin = np.array([0, len(w)], np.int64)
        rc = emu.emu_draw_commands(frames.ctypes.data_as(C.c_void_p), 1, 8, 8, w.ctypes.data_as(C.c_void_p),
                                   begin.ctypes.data_as(C.c_void_p), None)
        assert rc == -1, words
    assert not frames.any()


def test_polygons_with_more_edges_than_a_warp_orders_and_far_away_vertices():
    import cv2
    rng = np.random.default_rng(9)
    for t in range(30):
        w, h = int(rng.integers(40, 200)), int(rng.integers(40, 160))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        nv = int(rng.integers(130, 400))                      # > 128 edges: scan-converted by the host half
        pts = np.column_stack([rng.integers(-30, w + 30, nv), rng.integers(-30, h + 30, nv)]).astype(np.int32)
        far = np.array([[int(rng.integers(-10**6, 10**6)), int(rng.integers(-10**6, 10**6))] for _ in range(5)], np.int32)
        ref = img.copy()
        dl = DrawList(1)
        cv2.fillPoly(ref, [pts], (9, 8, 7))
        dl.fillPoly(0, pts, (9, 8, 7))
        cv2.fillPoly(ref, [far], (1, 2, 3))
        dl.fillPoly(0, far, (1, 2, 3))
        cv2.polylines(ref, [far], True, (4, 5, 6), 3)
        dl.polylines(0, far, True, (4, 5, 6), 3)
        cv2.line(ref, tuple(int(v) for v in far[0]), tuple(int(v) for v in far[1]), (7, 7, 7), 1)
        dl.line(0, far[0], far[1], (7, 7, 7), 1)
        cv2.circle(ref, (int(far[2][0]) % w, int(far[2][1]) % h), 30, (0, 9, 0), -1)
        dl.circle(0, (int(far[2][0]) % w, int(far[2][1]) % h), 30, (0, 9, 0), -1)
        mine = img.copy()[None]
        emu_commands(dl, mine)
        assert np.array_equal(ref, mine[0]), t
