from .lane_detector import LaneDetector, LaneLine

__all__ = ["LaneDetector", "LaneLine"]
