"""GPU-box probe: schedule of k4_ppht_v3 over the bench's 256-frame batch.

    make -C multimodal_autonomous_driving_perception_and_planning_b200/csrc prof          # -DLANE_PPHT_PROF build -> tools/_alt/
    LANE_B200_LIB=$PWD/tools/_alt/liblane_prof.so LANE_B200_PPHT_PROF=1 python tools/ppht_prof.py 256

Every frame's cluster records %globaltimer at its start and end, its SM and thread 0's cycle accounting in a global array;
a one-thread kernel prints the lines after the launch ("GT f=.. sm=.. start=.. end=.. vote=.. xchg=.. walk=.. total=..").
profiles/r2_ppht_schedule.txt is the output for the second (warm) call."""
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
