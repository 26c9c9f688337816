"""GPU-box probe: schedule of k4_ppht_v3 (a -DLANE_PPHT_PROF build selected with LANE_B200_LIB, run with
LANE_B200_PPHT_PROF=1): the bench's 256-frame batch, the kernel prints start / end (globaltimer) per frame."""
import sys
sys.path.insert(0, '.')
import torch
from multimodal_autonomous_driving_perception_and_planning_b200 import LaneDetector, SyntheticDataGenerator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = SyntheticDataGenerator(1920, 1080).generate_batch_device(64, start_frame=0)
frames = base.repeat((n + 63) // 64, 1, 1, 1)[:n].contiguous()
det = LaneDetector(max_batch=n)
det.detect_batch(frames)
torch.cuda.synchronize()
print('WARM', flush=True)
det.detect_batch(frames)
torch.cuda.synchronize()
