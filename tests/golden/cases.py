"""Seeded edge-case inputs shared by make_golden.py (reference side) and the tests.

`Gen` is a SyntheticDataGenerator class: the reference's bytecode class when generating the
fixtures, the re-created one (hash-checked against it) at test time.
"""
import cv2
import numpy as np


def build_cases(Gen):
    cases = {}
    rng = np.random.default_rng(1234)
    cases["noise_481x643"] = [rng.integers(0, 256, (481, 643, 3), dtype=np.uint8) for _ in range(2)]
    cases["black_480x640"] = [np.zeros((480, 640, 3), np.uint8)]
    g = Gen()
    one = [g.generate_frame_with_vehicles() for _ in range(3)]
    for f in one:
        f[:, 320:] = 60                                   # flatten the right half
    cases["onesided_480x640"] = one
    g = Gen()
    a = g.generate_frame_with_vehicles()
    cases["hit_miss_hit"] = [a, np.zeros_like(a), g.generate_frame_with_vehicles()]
    cases["tiny_7x9"] = [rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)]
    # median >= 196.5 => high == 255: strong pixels need |dx|+|dy| >= 256
    bright = np.clip(rng.normal(235, 6, (200, 320, 3)), 0, 255).astype(np.uint8)
    cv2.line(bright, (40, 190), (150, 20), (0, 0, 0), 3)
    cv2.line(bright, (300, 195), (170, 30), (10, 10, 10), 2)
    cases["bright_200x320"] = [bright]
    return cases


CUSTOM_ROI = np.array([[(0, 479), (0, 200), (639, 200), (639, 479)]], dtype=np.int32)


def custom_roi_frames(Gen):
    g = Gen()
    return [g.generate_frame_with_vehicles() for _ in range(2)]
