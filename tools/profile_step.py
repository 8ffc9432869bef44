"""One 1080p network call for ncu (eager launches, no graph): W warm-up calls + 1 profiled call."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frame_interpolation_b200 import synthetic
from frame_interpolation_b200.interpolator import Interpolator
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = Interpolator("synthetic", align=64)
eng.set_option("use_graph", 0)
x0, x1 = synthetic.frame_pair(1080, 1920, seed=0, n_waves=4)
d0, d1 = torch.from_numpy(x0).cuda(), torch.from_numpy(x1).cuda()
out = torch.empty_like(d0)
torch.cuda.synchronize()
for _ in range(warm + 1):
    eng.interpolate_device(d0.data_ptr(), d1.data_ptr(), 1, 1080, 1920, out.data_ptr())
    eng.synchronize()
print("done", eng.profile()["kernel_launches"])
