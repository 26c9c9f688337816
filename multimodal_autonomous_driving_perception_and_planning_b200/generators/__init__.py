from .synthetic_data import SyntheticDataGenerator, bgr_to_nv12, multi_camera_batch

__all__ = ["SyntheticDataGenerator", "bgr_to_nv12", "multi_camera_batch"]
