"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12)`` (uint8).

SURVEY.md section 8(f) rank 1, second half: frames enter the reference through ``VideoDataLoader.read_frame`` /
``read_frame_at`` (/root/reference/data/loaders/video_loader.py:96-131), i.e. out of a video decoder.  A hardware
decoder (NVDEC) hands out NV12 -- a full-resolution Y plane followed by a half-resolution interleaved UV plane,
1.5 B/px instead of 3 -- and the BGR frame the lane path consumes is what OpenCV's own NV12 -> BGR conversion makes
of it.  Moving that conversion to the device halves the host->device bytes per frame.

The arithmetic lives in OpenCV (opencv-python >= 4.5.0, installed 4.13.0.92; source not on this box); the
restatement follows the published algorithm (imgproc color_yuv: ITU-R BT.601 limited range, 20-bit fixed point) and
is pinned by tests/test_oracle_nv12.py against cv2 itself.

    layout  uint8 [H * 3 / 2][W]: rows 0..H-1 = Y; rows H.. = (U, V) byte pairs, one pair per 2x2 block of pixels
    u = U - 128, v = V - 128, y = max(0, Y - 16) * 1220542
    R = sat_u8((y + 2^19 + 1673527 v) >> 20)
    G = sat_u8((y + 2^19 -  852492 v - 409993 u) >> 20)
    B = sat_u8((y + 2^19 + 2116026 u) >> 20)              (arithmetic shift: the sums can be negative)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np

CY, CUB, CUG, CVG, CVR, SHIFT = 1220542, 2116026, -409993, -852492, 1673527, 20


def nv12_to_bgr(nv12: np.ndarray) -> np.ndarray:
    """uint8 [H*3/2, W] (H, W even) -> uint8 [H, W, 3] BGR."""
    nv12 = np.asarray(nv12, dtype=np.uint8)
    rows, w = nv12.shape
    if rows % 3 or w % 2 or (rows * 2 // 3) % 2:
        raise ValueError("NV12 needs even width and height")
    h = rows * 2 // 3
    y = np.maximum(0, nv12[:h].astype(np.int64) - 16) * CY + (1 << (SHIFT - 1))
    uv = nv12[h:].reshape(h // 2, w // 2, 2).astype(np.int64) - 128
    u = np.repeat(np.repeat(uv[..., 0], 2, axis=0), 2, axis=1)
    v = np.repeat(np.repeat(uv[..., 1], 2, axis=0), 2, axis=1)
    r = (y + CVR * v) >> SHIFT
    g = (y + CVG * v + CUG * u) >> SHIFT
    b = (y + CUB * u) >> SHIFT
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)


def bgr_to_nv12_for_tests(bgr: np.ndarray) -> np.ndarray:
    """A plausible NV12 image of a BGR frame (cv2's I420 conversion with the chroma planes interleaved).  Used only to
    make test / bench inputs that look like decoder output; nothing is pinned on this direction."""
    import cv2
    h, w = bgr.shape[:2]
    i420 = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420)
    out = np.empty((h * 3 // 2, w), np.uint8)
    out[:h] = i420[:h]
    u = i420[h:h + h // 4].reshape(h // 2, w // 2)
    v = i420[h + h // 4:].reshape(h // 2, w // 2)
    out[h:] = np.stack([u, v], axis=-1).reshape(h // 2, w)
    return out
