"""Oracle (both layers) against the golden fixtures produced by the unmodified reference
LaneDetector (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from cases import CUSTOM_ROI, build_cases, custom_roi_frames
from multimodal_autonomous_driving_perception_and_planning_b200.generators import SyntheticDataGenerator
from oracle import stages as S
from oracle.cv2_pipeline import Cv2LaneOracle
from util import gen_frames, golden, h16, lines_of, unpack_edges


def _check_stream(frames, gold, roi=None, every=1):
    so = S.StageOracle(roi_vertices=roi)
    co = Cv2LaneOracle(roi)
    for i, f in enumerate(frames):
        h, w = f.shape[:2]
        tr = so.step(f)
        lf, rf = co.detect(f)
        if i % every:
            continue
        assert tr.median_x2 == gold["median_x2"][i] and tr.low == gold["low"][i] and tr.high == gold["high"][i]
        assert h16(tr.blurred) == str(gold["blur_hash"][i])
        assert h16(tr.canny.edges) == str(gold["edge_hash"][i])
        assert int((tr.canny.edges != 0).sum()) == gold["n_edges"][i]
        assert int((tr.masked != 0).sum()) == gold["n_roi"][i]
        assert np.array_equal(tr.lines, lines_of(gold, i))
        for s, (a, b) in enumerate(((tr.left, lf), (tr.right, rf))):
            assert (a is not None) == bool(gold["valid"][i, s]) == (b is not None)
            if a is not None:
                for fit in (a, b):
                    assert np.array_equal(fit.coeffs, gold["poly"][i, s])
                    assert np.array_equal(fit.points, gold["points"][i, s])
                    assert fit.confidence == gold["conf"][i, s]
        off = gold["offset"][i]
        assert (tr.offset is None) == bool(np.isnan(off))
        if tr.offset is not None:
            assert tr.offset == off


def test_config1_300_frames():
    gold = golden("config1_640x480")
    _check_stream(gen_frames(640, 480, 300), gold)
    assert np.allclose(gold["offset"][:10], [7.5, 5.0, 3.5, 2.5, 1.0, -4.0, -3.5, -3.0, -3.0, -2.5])
    assert gold["valid"].all()


@pytest.mark.parametrize("name,w,h,n,start", [("hd1080_cam0", 1920, 1080, 4, 0), ("hd1080_cam1", 1920, 1080, 2, 1000),
                                               ("hd720", 1280, 720, 2, 0), ("uhd2160", 3840, 2160, 1, 0)])
def test_larger_resolutions(name, w, h, n, start):
    gold = golden(name)
    frames = gen_frames(w, h, n, start)
    _check_stream(frames, gold)
    so = S.StageOracle()
    for i, f in enumerate(frames):
        assert np.array_equal(so.step(f).canny.edges, unpack_edges(gold["edges_packed"][i], h, w))


def test_edge_cases():
    for name, frames in build_cases(SyntheticDataGenerator).items():
        gold = golden("case_" + name)
        assert h16(np.stack(frames)) == str(gold["frames_hash"]), name
        _check_stream(frames, gold)
    gold = golden("case_black_480x640")
    assert not gold["valid"].any() and np.isnan(gold["offset"]).all()
    gold = golden("case_onesided_480x640")
    assert gold["valid"][:, 0].all() and not gold["valid"][:, 1].any()


def test_custom_roi():
    gold = golden("case_custom_roi")
    assert np.array_equal(gold["roi"], CUSTOM_ROI)
    _check_stream(custom_roi_frames(SyntheticDataGenerator), gold, roi=CUSTOM_ROI)
