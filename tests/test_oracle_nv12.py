"""Pins oracle/nv12.py (the NV12 -> BGR restatement) against cv2 itself: random planes (every Y/U/V combination the
clamps can see), the limits, and generator frames taken through a BGR -> NV12 -> BGR round trip."""
import cv2
import numpy as np
import pytest

from oracle import nv12 as onv
from util import gen_frames


@pytest.mark.parametrize("h,w,seed", [(2, 2, 0), (6, 8, 1), (48, 64, 2), (480, 640, 3), (1080, 1920, 4), (34, 18, 5)])
def test_random_planes_match_cv2(h, w, seed):
    rng = np.random.default_rng(seed)
    nv12 = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
    assert np.array_equal(onv.nv12_to_bgr(nv12), cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12))


def test_extremes_and_every_luma_value():
    # all 256 luma values against the four chroma corners and the neutral point
    for u, v in [(0, 0), (0, 255), (255, 0), (255, 255), (128, 128), (16, 240)]:
        nv12 = np.empty((24, 256), np.uint8)
        nv12[:16] = np.arange(256, dtype=np.uint8)[None, :]
        nv12[16:, 0::2] = u
        nv12[16:, 1::2] = v
        assert np.array_equal(onv.nv12_to_bgr(nv12), cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12)), (u, v)


def test_generator_frames_through_nv12():
    for f in gen_frames(640, 480, 2):
        nv12 = onv.bgr_to_nv12_for_tests(f)
        assert nv12.shape == (720, 640)
        back = onv.nv12_to_bgr(nv12)
        assert np.array_equal(back, cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12))
        assert np.abs(back.astype(int) - f.astype(int)).mean() < 4      # a sane image, not a parity claim


def test_bad_shapes_raise():
    with pytest.raises(ValueError):
        onv.nv12_to_bgr(np.zeros((9, 7), np.uint8))
