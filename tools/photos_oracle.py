"""BASELINE.json configs[0] (plumbing): the reference's own fixture photos/one.png + two.png at t = 0.5
through the CPU oracle, run in the BUILD container (the photos live in /root/reference, which does not
exist on the GPU box). Writes timing + output statistics; the frames themselves are not copied into the repo."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frame_interpolation_b200 import eval_util, weights
from oracle.film_oracle import OracleInterpolator
root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/photos"
x0 = eval_util.read_image(os.path.join(root, "one.png"))[None]
x1 = eval_util.read_image(os.path.join(root, "two.png"))[None]
torch.set_num_threads(os.cpu_count())
orc = OracleInterpolator(weights.synthetic_weights(1234), align=64)
t = time.perf_counter()
mid = orc(x0, x1, np.full((1,), 0.5, np.float32))
el = time.perf_counter() - t
print(json.dumps({"config": "photos/one.png + two.png, t=0.5, CPU oracle (torch-CPU), synthetic Style weights",
                  "shape": list(x0.shape), "mean_abs_frame_difference": float(np.abs(x0 - x1).mean()),
                  "seconds": el, "threads": os.cpu_count(), "output_min": float(mid.min()), "output_max": float(mid.max()),
                  "output_mean": float(mid.mean())}, indent=1))
