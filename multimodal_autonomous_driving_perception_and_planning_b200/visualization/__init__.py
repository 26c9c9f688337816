from .draw_list import DrawList, color_word
from .overlays import OverlayRenderer, draw_lanes_batch

__all__ = ["DrawList", "color_word", "OverlayRenderer", "draw_lanes_batch"]
