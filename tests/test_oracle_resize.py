"""Pins oracle/resize.py (restatement of cv2.resize INTER_LINEAR, uint8) against cv2 itself (SURVEY 8f rank 1:
/root/reference/data/loaders/video_loader.py:108,128)."""
import cv2
import numpy as np
import pytest

from oracle import resize as orz

FIXED = [(1920, 1080, 1280, 720), (1920, 1080, 640, 480), (640, 480, 1920, 1080), (64, 48, 640, 480),
         (1000, 700, 333, 217), (333, 217, 1000, 700), (5, 7, 640, 480), (640, 480, 5, 7), (1, 1, 8, 8), (2, 3, 3, 2),
         (1920, 1080, 1919, 1079), (100, 100, 101, 99), (1280, 720, 1920, 1080), (1920, 1080, 960, 540)]


@pytest.mark.parametrize("sw,sh,dw,dh", FIXED)
@pytest.mark.parametrize("cn", [3, 1])
def test_fixed_sizes_match_cv2(sw, sh, dw, dh, cn):
    rng = np.random.default_rng(sw * 7 + dh)
    src = rng.integers(0, 256, (sh, sw, cn) if cn > 1 else (sh, sw), dtype=np.uint8)
    assert np.array_equal(orz.resize_linear(src, (dw, dh)), cv2.resize(src, (dw, dh)))


def test_random_sizes_match_cv2():
    rng = np.random.default_rng(11)
    for _ in range(60):
        sw, sh, dw, dh = (int(v) for v in rng.integers(1, 300, 4))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(orz.resize_linear(src, (dw, dh)), cv2.resize(src, (dw, dh))), (sw, sh, dw, dh)


def test_extreme_values_and_generator_frame():
    from multimodal_autonomous_driving_perception_and_planning_b200.generators.synthetic_data import multi_camera_batch
    f = multi_camera_batch(1, 1, 1920, 1080)[0][0]
    for dsize in [(640, 480), (1280, 720), (3840, 2160)]:
        assert np.array_equal(orz.resize_linear(f, dsize), cv2.resize(f, dsize))
    for v in (0, 255):
        src = np.full((37, 53, 3), v, np.uint8)
        assert np.array_equal(orz.resize_linear(src, (91, 17)), cv2.resize(src, (91, 17)))
    chk = np.indices((64, 64)).sum(0) % 2 * 255
    src = np.stack([chk] * 3, -1).astype(np.uint8)
    assert np.array_equal(orz.resize_linear(src, (100, 41)), cv2.resize(src, (100, 41)))
