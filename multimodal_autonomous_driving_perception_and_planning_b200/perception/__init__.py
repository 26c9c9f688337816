from .lane_detector import LaneDetector, LaneLine
from .road_layout import RoadLayoutAnalyzer, RoadLayoutCues

__all__ = ["LaneDetector", "LaneLine", "RoadLayoutAnalyzer", "RoadLayoutCues"]
