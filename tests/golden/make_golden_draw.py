"""Drawing goldens from the UNMODIFIED reference (run only where /root/reference is mounted):

    python tests/golden/make_golden_draw.py

Runs the reference's own ``LaneDetector.draw_lanes`` (src/perception/lane_detector.py:220-251) and
``OverlayRenderer.draw_lane_offset_indicator`` (src/visualization/overlays.py:103-148) on generator streams (lanes from
the reference's ``detect``) and on the random cases of draw_cases.py, and stores the sha256 of every annotated frame plus
the lane points / offsets that went in (draw_golden.npz), so that the tests can replay the same calls without the
reference tree.  cv2 / numpy as installed.
"""
import hashlib
import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
from draw_cases import N_RANDOM, STREAMS, random_case  # noqa: E402
from make_golden import load_generator  # noqa: E402
from src.perception.lane_detector import LaneDetector, LaneLine  # noqa: E402
from src.visualization.overlays import OverlayRenderer  # noqa: E402

warnings.simplefilter("ignore")


def h16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    Gen = load_generator()
    ov = OverlayRenderer()
    out = {}
    for w, h, n in STREAMS:
        gen = Gen(w, h)
        det = LaneDetector()
        pts = np.zeros((n, 2, 50, 2), np.int32)
        valid = np.zeros((n, 2), np.uint8)
        offs = np.full(n, np.nan)
        hashes = []
        for i, f in enumerate(gen.generate_video_stream(n)):
            l, r = det.detect(f)
            for s, lane in enumerate((l, r)):
                if lane is not None:
                    pts[i, s], valid[i, s] = lane.points, 1
            off = det.get_lane_center_offset(w, l, r)
            if off is not None:
                offs[i] = off
            filled = det.draw_lanes(f.copy(), l, r, True)
            lines_only = det.draw_lanes(f.copy(), l, r, False)
            both = ov.draw_lane_offset_indicator(filled.copy(), off)
            hashes.append([h16(filled), h16(lines_only), h16(both)])
        key = f"stream_{w}x{h}"
        out[key + "_points"], out[key + "_valid"], out[key + "_offset"] = pts, valid, offs
        out[key + "_hash"] = np.array(hashes)
    det = LaneDetector()
    hashes = []
    for seed in range(N_RANDOM):
        frame, pts, valid, off = random_case(seed)
        lanes = [LaneLine(points=pts[s], side="left", confidence=1.0) if valid[s] else None for s in range(2)]
        filled = det.draw_lanes(frame.copy(), lanes[0], lanes[1], True)
        lines_only = det.draw_lanes(frame.copy(), lanes[0], lanes[1], False)
        both = ov.draw_lane_offset_indicator(filled.copy(), off)
        hashes.append([h16(filled), h16(lines_only), h16(both)])
    out["random_hash"] = np.array(hashes)
    np.savez_compressed(os.path.join(HERE, "draw_golden.npz"), **out)
    print("wrote draw_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
