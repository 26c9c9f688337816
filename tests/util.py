"""Shared helpers for the test-suite (frames, golden loading)."""
import hashlib
import json
import os

import numpy as np

from multimodal_autonomous_driving_perception_and_planning_b200.generators import SyntheticDataGenerator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def h16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def meta():
    return json.load(open(os.path.join(GOLDEN, "meta.json")))


def gen_frames(w, h, n, start=0):
    g = SyntheticDataGenerator(w, h)
    g.frame_count = start
    return [g.generate_frame_with_vehicles() for _ in range(n)]


def lines_of(gold, i):
    o = gold["lines_off"]
    return gold["lines"][o[i]:o[i + 1]]


def unpack_edges(packed, h, w):
    return (np.unpackbits(packed)[: h * w].reshape(h, w) * 255).astype(np.uint8)
