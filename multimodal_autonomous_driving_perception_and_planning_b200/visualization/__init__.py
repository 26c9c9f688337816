from .draw_list import DrawList, color_word
from .overlays import OverlayRenderer, draw_lanes_batch, draw_lanes_records

__all__ = ["DrawList", "color_word", "OverlayRenderer", "draw_lanes_batch", "draw_lanes_records"]
