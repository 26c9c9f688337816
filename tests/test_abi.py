"""The C-ABI library loads on a CPU-only box, exports every symbol include/lane_b200.h declares,
agrees with the Python struct mirror, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from multimodal_autonomous_driving_perception_and_planning_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lane_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"LANE_API\s+[\w\s\*]+?\b(lane_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ("lane_ctx_create", "lane_ctx_destroy", "lane_set_roi_mask", "lane_set_threshold_lut",
                 "lane_detect_batch", "lane_detect_enqueue", "lane_detect_collect", "lane_debug_tap",
                 "lane_hough_accumulator", "lane_last_error"):
        assert must in names
    assert len(names) == len(_native.exported_symbols()) >= 22


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in lane_b200.h but not exported"
    assert set(declared_symbols()) == set(_native.exported_symbols())
    assert lib.lane_abi_version() == 1


def test_header_cites_the_reference_for_every_entry_point():
    text = open(HEADER).read()
    assert "lane_detector.py:178-218" in text and "lane_detector.py:253-272" in text
    assert text.count("lane_detector.py") >= 8


def test_struct_mirror_matches_header_layout():
    assert ctypes.sizeof(_native.LaneSide) == 4 + 4 + 24 + 24 + 8 + 50 * 2 * 4
    assert ctypes.sizeof(_native.LaneRecord) == _native.RECORD_DTYPE.itemsize
    rec = np.zeros(1, _native.RECORD_DTYPE)
    assert rec["side"]["points"].shape == (1, 2, 50, 2)
    for field in ("offset", "offset_valid", "median_x2", "low", "high", "n_edges", "n_roi_points", "n_segments",
                  "hysteresis_rounds", "flags", "n_segments_found"):
        assert _native.RECORD_DTYPE.fields[field][1] == getattr(_native.LaneRecord, field).offset


def test_threshold_lut_is_the_reference_expression():
    low, high = _native.threshold_lut()
    for k in range(511):
        m = k / 2.0
        assert low[k] == int(max(0, 0.7 * m)) and high[k] == int(min(255, 1.3 * m))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = _native.lib().lane_ctx_create(0, 480, 640, 4, 0, ctypes.byref(h))
    assert rc == -3 and not h.value
    assert b"no CPU fallback" in _native.lib().lane_last_error(None)
    from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector
    with pytest.raises(_native.LaneError):
        LaneDetector().detect(np.zeros((48, 64, 3), np.uint8))


def test_bad_arguments_are_rejected_without_a_device():
    h = ctypes.c_void_p()
    lib = _native.lib()
    assert lib.lane_ctx_create(0, 0, 640, 4, 0, ctypes.byref(h)) == -1
    assert lib.lane_ctx_create(0, 480, 40000, 4, 0, ctypes.byref(h)) == -5
    assert lib.lane_set_roi_mask(None, None) == -1
    assert lib.lane_detect_batch(None, None, 0, 1, None, 1, None, None, None) == -1
