"""Host-side mirror of the reference interface: input checks, ROI rasterisation, record decoding,
drawing -- everything that needs no GPU."""
import cv2
import numpy as np
import pytest

from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, LaneLine, _native
from oracle import stages as S


def test_constructor_matches_reference_signature():
    d = LaneDetector()
    assert d.roi_vertices is None and d.prev_left_fit is None and d.prev_right_fit is None
    assert d.smoothing_factor == 0.7
    roi = np.array([[(0, 10), (5, 0), (10, 10)]], np.int32)
    assert LaneDetector(roi).roi_vertices is roi
    line = LaneLine(points=np.zeros((50, 2), np.int32), side="left", confidence=0.5)
    assert line.polynomial is None


def test_input_errors_are_cv2_errors_like_the_reference():
    d = LaneDetector()
    for bad in (np.zeros((48, 64, 3), np.float32), np.zeros((48, 64), np.uint8), np.zeros((48, 64, 4), np.uint8)):
        with pytest.raises(cv2.error):
            d.detect(bad)
    with pytest.raises(cv2.error):
        d.detect_batch(np.zeros((2, 48, 64), np.uint8))
    assert d.detect_batch(np.zeros((0, 48, 64, 3), np.uint8)) == []


@pytest.mark.parametrize("shape", [(480, 640), (1080, 1920), (7, 9), (481, 643)])
def test_roi_mask_equals_reference_rasterisation(shape):
    assert np.array_equal(LaneDetector()._get_roi_mask(shape), S.roi_mask(*shape))


def test_offset_and_reset():
    d = LaneDetector()
    pts = np.zeros((50, 2), np.int32)
    left = LaneLine(points=pts.copy(), side="left", confidence=1.0)
    right = LaneLine(points=pts.copy(), side="right", confidence=1.0)
    left.points[-1, 0], right.points[-1, 0] = 176, 452
    assert d.get_lane_center_offset(640, left, right) == 640 / 2 - (176 + 452) / 2
    assert d.get_lane_center_offset(640, None, right) is None
    d.prev_left_fit = np.ones(3)
    d.reset()
    assert d.prev_left_fit is None and d.prev_right_fit is None


def test_records_decode_to_lane_lines():
    d = LaneDetector()
    recs = np.zeros(2, _native.RECORD_DTYPE)
    recs[0]["side"][0]["valid"] = 1
    recs[0]["side"][0]["coeffs"] = [1.0, 2.0, 3.0]
    recs[0]["side"][0]["confidence"] = 0.3
    recs[0]["side"][0]["points"][:, 0] = np.arange(50)
    out = d._lanes_from_records(recs)
    assert out[1] == (None, None) and out[0][1] is None
    lane = out[0][0]
    assert lane.side == "left" and lane.confidence == 0.3 and lane.points.dtype == np.int32
    assert np.array_equal(lane.polynomial, [1.0, 2.0, 3.0]) and lane.points.shape == (50, 2)


def test_draw_lanes_fill_and_in_place_behaviour():
    d = LaneDetector()
    frame = np.full((100, 120, 3), 50, np.uint8)
    ys = np.linspace(60, 100, 50).astype(np.int32)
    left = LaneLine(points=np.stack([np.full(50, 30), ys], 1).astype(np.int32), side="left", confidence=1.0)
    right = LaneLine(points=np.stack([np.full(50, 90), ys], 1).astype(np.int32), side="right", confidence=1.0)
    out = d.draw_lanes(frame, left, right)
    assert out is not frame and (frame == 50).all()            # filled overlay comes back as a new image
    assert tuple(out[80, 60]) == (35, 112, 65)                  # 0.7*50 + 0.3*(0,255,100)
    assert tuple(out[80, 30]) == (255, 0, 0) and tuple(out[80, 90]) == (0, 0, 255)
    only_left = d.draw_lanes(frame, left, None)
    assert only_left is frame and tuple(frame[80, 30]) == (255, 0, 0)   # the reference draws in place here


def test_frame_ingest_and_scene_stats_validate_inputs_before_touching_the_gpu():
    import cv2
    import pytest
    from multimodal_autonomous_driving_perception_and_planning_b200 import FrameIngest
    from multimodal_autonomous_driving_perception_and_planning_b200.perception.scene_stats import SceneStatsAnalyzer
    frames = np.zeros((2, 8, 8, 3), np.uint8)
    assert FrameIngest(None).resize_batch(frames) is frames                      # target_size=None: no resize, as the loader
    assert FrameIngest((4, 4)).target_size == (4, 4)
    with pytest.raises(cv2.error):
        FrameIngest((4, 4)).resize_batch(frames.astype(np.float32))
    with pytest.raises(cv2.error):
        FrameIngest((4, 4)).resize_batch(np.zeros((2, 8, 8, 2), np.uint8))
    with pytest.raises(ValueError):
        SceneStatsAnalyzer().analyze_batch(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(ValueError):
        SceneStatsAnalyzer().analyze_batch(frames.astype(np.int16))
