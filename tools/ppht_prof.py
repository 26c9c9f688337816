import numpy as np, sys
sys.path.insert(0,'/root/repo')
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, multi_camera_batch
b=multi_camera_batch(1,64,1920,1080,period=8)[0]
det=LaneDetector(max_batch=64)
det.detect_batch(b)
import torch; torch.cuda.synchronize()
print('segments', det.last_records['n_segments'][:8], 'roi pts', det.last_records['n_roi_points'][:8])
