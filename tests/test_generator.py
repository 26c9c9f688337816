"""The re-created generator must reproduce the reference bytecode's frames (sha256 fixtures
recorded by tests/golden/make_golden.py from the .pyc; SURVEY.md section 8c lists the same values)."""
import numpy as np

from multimodal_autonomous_driving_perception_and_planning_b200.generators import (
    SyntheticDataGenerator, multi_camera_batch)
from util import h16, meta


def test_frame_hashes_match_reference_bytecode():
    want = meta()["frame_hashes"]
    by_res = {}
    for key in want:
        if key.endswith("all300"):
            continue
        res, idx = key.split(":")
        by_res.setdefault(res, []).append(int(idx))
    for res, idxs in by_res.items():
        w, h = map(int, res.split("x"))
        for idx in idxs:
            g = SyntheticDataGenerator(w, h)
            if idx >= 1000:
                g.frame_count = idx
                f = g.generate_frame_with_vehicles()
            else:
                for _ in range(idx + 1):
                    f = g.generate_frame_with_vehicles()
            assert h16(f) == want[f"{res}:{idx}"], key


def test_default_stream_is_300_frames_640x480():
    g = SyntheticDataGenerator()
    frames = list(g.generate_video_stream())
    assert len(frames) == 300 and frames[0].shape == (480, 640, 3) and frames[0].dtype == np.uint8
    assert h16(np.stack(frames)) == meta()["frame_hashes"]["640x480:all300"]


def test_generator_does_not_touch_global_rng():
    np.random.seed(7)
    a = np.random.rand()
    np.random.seed(7)
    SyntheticDataGenerator().generate_frame_with_vehicles()
    assert np.random.rand() == a


def test_multi_camera_batch_layout():
    b = multi_camera_batch(3, 4, 320, 240, period=2)
    assert b.shape == (3, 4, 240, 320, 3)
    g = SyntheticDataGenerator(320, 240)
    g.frame_count = 2000
    assert np.array_equal(b[2, 0], g.generate_frame_with_vehicles())
    assert np.array_equal(b[1, 2], b[1, 0]) and not np.array_equal(b[0, 0], b[1, 0])
