from .synthetic_data import SyntheticDataGenerator, multi_camera_batch

__all__ = ["SyntheticDataGenerator", "multi_camera_batch"]
