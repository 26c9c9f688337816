"""Generate the golden fixtures from the UNMODIFIED reference.

Run only where /root/reference is mounted (the build container):

    python tests/golden/make_golden.py

It imports the reference's own LaneDetector (src/perception/lane_detector.py) and loads the
bytecode-only SyntheticDataGenerator (data/generators/__pycache__/synthetic_data.cpython-312.pyc),
runs them on cv2/numpy as installed, and writes small .npz/.json fixtures next to this file.
Nothing at test time reads /root/reference: the tests regenerate the frames with the
re-created generator (checked against the sha256 values stored here) and compare against
these stored outputs.
"""
import hashlib
import json
import marshal
import os
import sys
import types
import warnings

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
from cases import CUSTOM_ROI, build_cases, custom_roi_frames  # noqa: E402
from src.perception.lane_detector import LaneDetector  # noqa: E402

warnings.simplefilter("ignore")


def load_generator():
    p = os.path.join(REF, "data/generators/__pycache__/synthetic_data.cpython-312.pyc")
    code = marshal.loads(open(p, "rb").read()[16:])
    m = types.ModuleType("ref_synthetic_data")
    exec(code, m.__dict__)
    return m.SyntheticDataGenerator


def h16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run_stream(frames, roi=None):
    """Run a fresh reference detector over frames; collect every observable."""
    ld = LaneDetector(roi)
    n = len(frames)
    rec = dict(median_x2=np.zeros(n, np.int32), low=np.zeros(n, np.int32), high=np.zeros(n, np.int32),
               n_edges=np.zeros(n, np.int32), n_roi=np.zeros(n, np.int32),
               valid=np.zeros((n, 2), np.uint8), poly=np.zeros((n, 2, 3), np.float64),
               conf=np.zeros((n, 2), np.float64), points=np.zeros((n, 2, 50, 2), np.int32),
               offset=np.full(n, np.nan, np.float64))
    edge_hash, lines_all, lines_off, edges_list, blur_hash = [], [], [0], [], []
    for i, f in enumerate(frames):
        h, w = f.shape[:2]
        pre = ld._preprocess(f)
        m = np.median(pre)
        rec["median_x2"][i] = int(round(2 * m))
        rec["low"][i] = int(max(0, 0.7 * m))
        rec["high"][i] = int(min(255, 1.3 * m))
        e = ld._detect_edges(pre)
        me = ld._apply_roi(e)
        ln = ld._detect_lines(me)
        ln = np.zeros((0, 4), np.int32) if len(ln) == 0 else np.asarray(ln, np.int32).reshape(-1, 4)
        rec["n_edges"][i] = int((e != 0).sum())
        rec["n_roi"][i] = int((me != 0).sum())
        blur_hash.append(h16(pre))
        edge_hash.append(h16(e))
        edges_list.append(e)
        lines_all.append(ln)
        lines_off.append(lines_off[-1] + len(ln))
        left, right = ld.detect(f)            # the public call (state advances here only)
        for s, lane in enumerate((left, right)):
            if lane is not None:
                rec["valid"][i, s] = 1
                rec["poly"][i, s] = lane.polynomial
                rec["conf"][i, s] = lane.confidence
                rec["points"][i, s] = lane.points
        off = ld.get_lane_center_offset(w, left, right)
        if off is not None:
            rec["offset"][i] = off
    rec["lines"] = np.concatenate(lines_all, 0) if lines_all else np.zeros((0, 4), np.int32)
    rec["lines_off"] = np.asarray(lines_off, np.int64)
    rec["edge_hash"] = np.asarray(edge_hash)
    rec["blur_hash"] = np.asarray(blur_hash)
    return rec, edges_list


def main():
    Gen = load_generator()
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "frame_hashes": {}}

    # generator hashes at every BASELINE resolution
    for (w, h, idxs) in [(640, 480, [0, 1, 2, 150, 299]), (1280, 720, [0, 1]), (1920, 1080, [0, 1, 2]),
                         (3840, 2160, [0])]:
        g = Gen(w, h)
        want = set(idxs)
        for i in range(max(idxs) + 1):
            f = g.generate_frame_with_vehicles()
            if i in want:
                meta["frame_hashes"][f"{w}x{h}:{i}"] = h16(f)
    g = Gen(1920, 1080)
    g.frame_count = 1000
    meta["frame_hashes"]["1920x1080:1000"] = h16(g.generate_frame_with_vehicles())

    # config 1: 300 frames 640x480 through a fresh reference detector
    g = Gen()
    frames = [g.generate_frame_with_vehicles() for _ in range(300)]
    meta["frame_hashes"]["640x480:all300"] = h16(np.stack(frames))
    rec, _ = run_stream(frames)
    np.savez_compressed(os.path.join(HERE, "config1_640x480.npz"), **rec)

    # 1080p: frames 0..3 of camera 0 and frames 0..1 of camera 1 (frame_count 1000..)
    g = Gen(1920, 1080)
    frames = [g.generate_frame_with_vehicles() for _ in range(4)]
    rec, edges = run_stream(frames)
    rec["edges_packed"] = np.stack([np.packbits(e != 0) for e in edges])
    np.savez_compressed(os.path.join(HERE, "hd1080_cam0.npz"), **rec)
    g = Gen(1920, 1080)
    g.frame_count = 1000
    frames = [g.generate_frame_with_vehicles() for _ in range(2)]
    rec, edges = run_stream(frames)
    rec["edges_packed"] = np.stack([np.packbits(e != 0) for e in edges])
    np.savez_compressed(os.path.join(HERE, "hd1080_cam1.npz"), **rec)

    # 720p and 4K single frames
    for (w, h, name) in [(1280, 720, "hd720"), (3840, 2160, "uhd2160")]:
        g = Gen(w, h)
        frames = [g.generate_frame_with_vehicles() for _ in range(2 if w < 3000 else 1)]
        rec, edges = run_stream(frames)
        rec["edges_packed"] = np.stack([np.packbits(e != 0) for e in edges])
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **rec)

    # edge cases: inputs are seeded so tests can rebuild them
    cases = build_cases(Gen)
    for name, frames in cases.items():
        rec, edges = run_stream(frames)
        rec["edges_packed"] = np.stack([np.packbits(e != 0) for e in edges])
        rec["frames_hash"] = np.asarray(h16(np.stack(frames)))
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **rec)
    # custom ROI (full-frame polygon) on a generator frame
    frames, roi = custom_roi_frames(Gen), CUSTOM_ROI
    rec, edges = run_stream(frames, roi)
    rec["roi"] = roi
    np.savez_compressed(os.path.join(HERE, "case_custom_roi.npz"), **rec)

    json.dump(meta, open(os.path.join(HERE, "meta.json"), "w"), indent=1, sort_keys=True)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
