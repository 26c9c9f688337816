"""GPU parity of K7 (csrc/k7_draw.cu), the batched cv2-exact rasteriser, through the C ABI (lane_draw_commands,
lane_draw_lanes_batch): random cv2 call mixes against cv2 itself, draw_lanes / the offset indicator against goldens made
by the unmodified reference, generator frames rasterised on the device against the host generator, and full-size
batches (config 2's 256 x 1080p) through size-independent properties."""
import hashlib

import cv2
import numpy as np
import pytest

from draw_cases import N_RANDOM, STREAMS, random_case
from draw_util import cv2_draw_lanes, cv2_offset_indicator, draw_golden, h16, random_mix
from multimodal_autonomous_driving_perception_and_planning_b200 import (DrawList, LaneDetector, OverlayRenderer,
                                                                        SyntheticDataGenerator, draw_lanes_batch)
from multimodal_autonomous_driving_perception_and_planning_b200.visualization.overlays import draw_lanes_arrays
from util import gen_frames

pytestmark = pytest.mark.gpu


def test_random_command_mixes_equal_cv2_on_the_device():
    rng = np.random.default_rng(10)
    for t in range(400):
        img, ref, dl = random_mix(rng, lambda: DrawList(1))
        mine = img.copy()[None]
        dl.execute(mine)
        assert np.array_equal(ref, mine[0]), t


def test_many_frames_with_different_lists_in_one_launch():
    import torch
    rng = np.random.default_rng(11)
    n, h, w = 64, 97, 131
    start = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    ref = start.copy()
    dl = DrawList(n)
    for f in range(n):
        for _ in range(int(rng.integers(0, 6))):
            col = tuple(int(c) for c in rng.integers(0, 256, 3))
            a = (int(rng.integers(-30, w + 30)), int(rng.integers(-30, h + 30)))
            b = (int(rng.integers(-30, w + 30)), int(rng.integers(-30, h + 30)))
            kind = int(rng.integers(0, 4))
            if kind == 0:
                th = int(rng.integers(1, 5))
                cv2.line(ref[f], a, b, col, th)
                dl.line(f, a, b, col, th)
            elif kind == 1:
                cv2.rectangle(ref[f], a, b, col, -1)
                dl.rectangle(f, a, b, col, -1)
            elif kind == 2:
                r = int(rng.integers(0, 30))
                cv2.circle(ref[f], a, r, col, -1)
                dl.circle(f, a, r, col, -1)
            else:
                pts = np.array([a, b, (int(rng.integers(-30, w + 30)), int(rng.integers(-30, h + 30)))], np.int32)
                o = ref[f].copy()
                cv2.fillPoly(o, [pts], col)
                ref[f] = cv2.addWeighted(ref[f], 0.7, o, 0.3, 0)
                dl.fillPoly_weighted(f, pts, col, 0.7, 0.3)
    dev = torch.from_numpy(start).cuda()
    out = dl.execute(dev)
    assert out.data_ptr() == dev.data_ptr()
    assert np.array_equal(ref, dev.cpu().numpy())


def test_draw_lanes_and_offset_indicator_equal_the_reference_goldens():
    import torch
    g = draw_golden()
    ov = OverlayRenderer()
    for w, h, n in STREAMS:
        key = f"stream_{w}x{h}"
        frames = np.stack(gen_frames(w, h, n))
        pts, valid = g[key + "_points"], g[key + "_valid"]
        for fill, col in ((True, 0), (False, 1)):
            dev = torch.from_numpy(frames).cuda()
            draw_lanes_arrays(dev, pts[:, 0], valid[:, 0], pts[:, 1], valid[:, 1], fill)
            assert [h16(f) for f in dev.cpu().numpy()] == [r[col] for r in g[key + "_hash"]]
            if fill:
                offs = [None if np.isnan(o) else float(o) for o in g[key + "_offset"]]
                ov.draw_lane_offset_indicator_batch(dev, offs)
                assert [h16(f) for f in dev.cpu().numpy()] == [r[2] for r in g[key + "_hash"]]
    for seed in range(N_RANDOM):
        frame, pts, valid, off = random_case(seed)
        for fill, col in ((True, 0), (False, 1)):
            mine = frame.copy()[None]
            draw_lanes_arrays(mine, pts[None, 0], valid[None, 0], pts[None, 1], valid[None, 1], fill)
            assert h16(mine[0]) == g["random_hash"][seed][col], (seed, fill)
            if fill:
                ov.draw_lane_offset_indicator_batch(mine, [off])
                assert h16(mine[0]) == g["random_hash"][seed][2], seed


def test_detect_then_draw_on_the_device_equals_the_cv2_sequence():
    import torch
    w, h, n = 640, 480, 24
    frames = np.stack(gen_frames(w, h, n))
    det = LaneDetector()
    dev = torch.from_numpy(frames).cuda()
    lanes = det.detect_batch(dev)
    out = det.draw_lanes_batch(dev.clone(), lanes)
    offs = [det.get_lane_center_offset(w, l, r) for l, r in lanes]
    OverlayRenderer().draw_lane_offset_indicator_batch(out, offs)
    got = out.cpu().numpy()
    for i, (l, r) in enumerate(lanes):
        ref = cv2_draw_lanes(frames[i].copy(), None if l is None else l.points, None if r is None else r.points)
        ref = cv2_offset_indicator(ref, offs[i])
        assert np.array_equal(ref, got[i]), i
    # the single-frame host method (cv2, as in the reference) and the batch method agree
    l, r = lanes[3]
    assert np.array_equal(det.draw_lanes(frames[3].copy(), l, r), cv2_draw_lanes(frames[3].copy(), l.points, r.points))


def test_degenerate_and_far_outside_lanes():
    rng = np.random.default_rng(12)
    n, h, w = 48, 173, 301
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    pts = np.zeros((n, 2, 50, 2), np.int32)
    valid = np.ones((n, 2), np.uint8)
    y = np.linspace(0.6 * h, h, 50)
    for t in range(n):
        for s in range(2):
            kind = t % 6
            if kind == 0:
                x = np.full(50, w // 2)
            elif kind == 1:
                x = np.full(50, (-3000 if s == 0 else 4000)) + rng.integers(-50, 50, 50)
            elif kind == 2:
                x = np.linspace(0, w, 50) if s == 0 else np.linspace(w, 0, 50)
            elif kind == 3:
                x = np.polyval([rng.normal(0, 0.05), rng.normal(0, 3), rng.normal(w / 2, w)], y)
            elif kind == 4:
                x = np.polyval([rng.normal(0, 1e-3), rng.normal(0, .5), rng.normal(w / 2, w / 3)], y)
                valid[t, s] = (t // 6 + s) % 2
            else:
                x = rng.integers(-40, w + 40, 50)          # zig-zag: many crossings per row
            pts[t, s] = np.column_stack([x, y]).astype(np.int32)
    got = frames.copy()
    draw_lanes_arrays(got, pts[:, 0], valid[:, 0], pts[:, 1], valid[:, 1], True)
    for t in range(n):
        ref = cv2_draw_lanes(frames[t].copy(), pts[t, 0] if valid[t, 0] else None, pts[t, 1] if valid[t, 1] else None)
        assert np.array_equal(ref, got[t]), t


def test_generator_on_the_device_equals_the_host_generator():
    for w, h, n, start in [(640, 480, 120, 0), (1920, 1080, 6, 95), (1280, 720, 4, 1000), (3840, 2160, 2, 3), (321, 203, 105, 7)]:
        ref = SyntheticDataGenerator(w, h).generate_batch(n, start_frame=start)
        for recorded in (False, True):       # scene laid out by the library's C++ code / recorded by the Python class
            gen = SyntheticDataGenerator(w, h)
            dev = gen.generate_batch_device(n, start_frame=start, recorded=recorded)
            assert gen.frame_count == start + n
            assert np.array_equal(ref, dev.cpu().numpy()), (w, h, recorded)
    # host buffers through the raw ABI
    import ctypes as C
    from multimodal_autonomous_driving_perception_and_planning_b200 import _native
    out = np.full((3, 120, 160, 3), 7, np.uint8)
    assert _native.lib().lane_generate_frames(out.ctypes.data_as(C.c_void_p), 0, 3, 120, 160, 41, 0, None, None) == 0
    assert np.array_equal(out, SyntheticDataGenerator(160, 120).generate_batch(3, start_frame=41))


def test_generator_fixtures_of_the_survey_on_the_device():
    # sha256[:16] of the reference generator's frames
    gen = SyntheticDataGenerator(640, 480)
    got = gen.generate_batch_device(300, start_frame=0).cpu().numpy()
    assert hashlib.sha256(got[0].tobytes()).hexdigest()[:16] == "7fbddeb953012a34"
    assert hashlib.sha256(got[150].tobytes()).hexdigest()[:16] == "05fa76424e2f691e"
    assert hashlib.sha256(got[299].tobytes()).hexdigest()[:16] == "aadaa7dabc6ad28b"
    assert hashlib.sha256(got.tobytes()).hexdigest()[:16] == "701c6dd0c4e8d707"
    assert hashlib.sha256(SyntheticDataGenerator(1920, 1080).generate_batch_device(1, 0).cpu().numpy().tobytes()
                          ).hexdigest()[:16] == "14758bdd16a47389"


def test_full_size_batch_properties():
    """Config 2's size (256 x 1080p): frames tiled from 8 distinct ones must come out as 8 distinct annotated frames
    tiled the same way (each checked against cv2), drawing twice the line-only overlay is idempotent, and a batch with no
    valid lane is untouched."""
    import torch
    w, h, n, distinct = 1920, 1080, 256, 8
    gen = SyntheticDataGenerator(w, h)
    base = gen.generate_batch_device(distinct, start_frame=0)
    dev = base.repeat(n // distinct, 1, 1, 1).contiguous()
    det = LaneDetector()
    lanes8 = det.detect_batch(base)
    lanes = lanes8 * (n // distinct)
    before = dev.clone()
    draw_lanes_batch(dev, lanes, True)
    host8 = base.cpu().numpy()
    for i in range(distinct):
        l, r = lanes8[i]
        ref = cv2_draw_lanes(host8[i].copy(), None if l is None else l.points, None if r is None else r.points)
        assert np.array_equal(ref, dev[i].cpu().numpy()), i
    assert torch.equal(dev.view(n // distinct, distinct, h, w, 3), dev[:distinct].unsqueeze(0).expand(n // distinct, -1, -1, -1, -1))
    lines = before.clone()
    draw_lanes_batch(lines, lanes, False)
    again = lines.clone()
    draw_lanes_batch(again, lanes, False)
    assert torch.equal(lines, again)
    untouched = before.clone()
    draw_lanes_batch(untouched, [(None, None)] * n, True)
    assert torch.equal(untouched, before)


def test_bad_arguments():
    import torch
    from multimodal_autonomous_driving_perception_and_planning_b200 import _native
    with pytest.raises(ValueError):
        DrawList(2).execute(np.zeros((1, 8, 8, 3), np.uint8))
    dl = DrawList(1)
    dl.extend(0, np.array([99, 1, 2], np.int32))
    with pytest.raises(_native.LaneError):
        dl.execute(np.zeros((1, 8, 8, 3), np.uint8))
    dl = DrawList(1)
    dl.fillPoly_weighted(0, [(0, 0), (4, 0), (4, 4)], (1, 2, 3), 0.3, 0.6, 4.0)      # gamma != 0 is refused
    with pytest.raises(_native.LaneError):
        dl.execute(np.zeros((1, 8, 8, 3), np.uint8))
    with pytest.raises(ValueError):
        draw_lanes_batch(torch.zeros((2, 8, 8, 3), dtype=torch.uint8, device="cuda"), [(None, None)])


def test_polygons_with_more_edges_than_a_warp_orders_and_far_away_vertices():
    rng = np.random.default_rng(9)
    for t in range(12):
        w, h = int(rng.integers(40, 200)), int(rng.integers(40, 160))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        nv = int(rng.integers(100, 400))
        pts = np.column_stack([rng.integers(-30, w + 30, nv), rng.integers(-30, h + 30, nv)]).astype(np.int32)
        far = np.array([[int(rng.integers(-10**6, 10**6)), int(rng.integers(-10**6, 10**6))] for _ in range(5)], np.int32)
        ref = img.copy()
        dl = DrawList(1)
        cv2.fillPoly(ref, [pts], (9, 8, 7))
        dl.fillPoly(0, pts, (9, 8, 7))
        o = ref.copy()
        cv2.fillPoly(o, [pts[:120]], (50, 60, 70))              # 120 edges: ordered per row by a warp, many crossings
        ref = cv2.addWeighted(ref, 0.7, o, 0.3, 0)
        dl.fillPoly_weighted(0, pts[:120], (50, 60, 70), 0.7, 0.3)
        cv2.fillPoly(ref, [far], (1, 2, 3))
        dl.fillPoly(0, far, (1, 2, 3))
        cv2.polylines(ref, [far], True, (4, 5, 6), 3)
        dl.polylines(0, far, True, (4, 5, 6), 3)
        mine = img.copy()[None]
        dl.execute(mine)
        assert np.array_equal(ref, mine[0]), t


def test_lane_overlay_kernel_and_primitive_lists_agree_and_records_on_the_device():
    """draw_lanes_batch runs k7_lanes (geometry on the device) by default; LANE_B200_DRAW_LANES=prims takes the host-expanded
    primitive lists.  Both must give cv2's pixels; and the overlay can be drawn straight from the device records."""
    import os
    import subprocess
    import sys
    import torch
    from multimodal_autonomous_driving_perception_and_planning_b200.visualization import draw_lanes_records
    w, h, n = 1280, 720, 40
    frames = SyntheticDataGenerator(w, h).generate_batch_device(n, start_frame=0)
    det = LaneDetector(max_batch=n)
    lanes = det.detect_batch(frames)
    by_kernel = det.draw_lanes_batch(frames.clone(), lanes)
    host = frames.cpu().numpy()
    for i, (l, r) in enumerate(lanes):
        ref = cv2_draw_lanes(host[i].copy(), None if l is None else l.points, None if r is None else r.points)
        assert np.array_equal(ref, by_kernel[i].cpu().numpy()), i
    from_records = draw_lanes_records(frames.clone(), det._ctx.records_device_ptr())
    torch.cuda.synchronize()
    assert torch.equal(from_records, by_kernel)
    no_fill = draw_lanes_records(frames.clone(), det._ctx.records_device_ptr(), fill_lane=False)
    assert torch.equal(no_fill, det.draw_lanes_batch(frames.clone(), lanes, fill_lane=False))
    # the primitive-list path in a child process (the switch is read once per process)
    code = ("import sys, numpy as np, torch; sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')\n"
            "from draw_cases import N_RANDOM, random_case\nfrom draw_util import draw_golden, h16\n"
            "from multimodal_autonomous_driving_perception_and_planning_b200.visualization.overlays import draw_lanes_arrays\n"
            "g = draw_golden()\n"
            "for seed in range(N_RANDOM):\n"
            "    frame, pts, valid, off = random_case(seed)\n"
            "    for fill, col in ((True, 0), (False, 1)):\n"
            "        mine = frame.copy()[None]\n"
            "        draw_lanes_arrays(mine, pts[None, 0], valid[None, 0], pts[None, 1], valid[None, 1], fill)\n"
            "        assert h16(mine[0]) == g['random_hash'][seed][col], (seed, fill)\n"
            "print('prims ok')\n")
    env = dict(os.environ, LANE_B200_DRAW_LANES="prims")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and "prims ok" in out.stdout, out.stderr[-2000:]
    det.close()
