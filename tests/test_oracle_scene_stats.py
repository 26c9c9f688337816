"""Pins oracle/scene_stats.py against cv2 / numpy (the SceneClassifier's image statistics,
/root/reference/src/tagging/scene_classifier.py:183-186, :237-238, :254)."""
import cv2
import numpy as np
import pytest

from oracle import scene_stats as oss


def _frames():
    from multimodal_autonomous_driving_perception_and_planning_b200.generators.synthetic_data import SyntheticDataGenerator
    rng = np.random.default_rng(3)
    out = [SyntheticDataGenerator(640, 480).generate_batch(2, start_frame=17)[1],
           rng.integers(0, 256, (97, 131, 3), dtype=np.uint8),
           rng.integers(0, 256, (1, 9, 3), dtype=np.uint8),
           rng.integers(0, 256, (9, 1, 3), dtype=np.uint8),
           np.zeros((20, 30, 3), np.uint8)]
    grass = np.zeros((64, 64, 3), np.uint8)
    grass[...] = (40, 160, 60)
    grass[::3] = (90, 200, 30)
    out.append(grass)
    return out


def test_hsv_matches_cv2_on_a_colour_lattice():
    lat = np.stack(np.meshgrid(*[np.arange(0, 256, 5)] * 3, indexing="ij"), -1).reshape(-1, 1, 3).astype(np.uint8)
    assert np.array_equal(oss.bgr2hsv(lat), cv2.cvtColor(lat, cv2.COLOR_BGR2HSV))
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    assert np.array_equal(oss.bgr2hsv(img), cv2.cvtColor(img, cv2.COLOR_BGR2HSV))


@pytest.mark.parametrize("idx", range(6))
def test_sums_and_cues_match_the_reference_expressions(idx):
    f = _frames()[idx]
    s = oss.frame_sums(f)
    gray = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    lap = cv2.Laplacian(gray, cv2.CV_64F)
    hsv = cv2.cvtColor(f, cv2.COLOR_BGR2HSV)
    green = cv2.inRange(hsv, (35, 40, 40), (85, 255, 255))
    assert s.sum_gray == int(gray.astype(np.int64).sum())
    assert np.array_equal(oss.laplacian(gray).astype(np.float64), lap)
    assert s.green_pixels == int(np.sum(green > 0))
    mean, var, ratio = oss.cues_from_sums(s)
    assert mean == np.mean(gray)                                   # exact: integer sum / N in float64
    assert ratio == np.sum(green > 0) / green.size
    assert var == pytest.approx(lap.var(), rel=1e-12, abs=1e-12)   # numpy's two-pass float64 variance
