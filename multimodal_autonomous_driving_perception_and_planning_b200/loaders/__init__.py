from .frame_ingest import FrameIngest

__all__ = ["FrameIngest"]
