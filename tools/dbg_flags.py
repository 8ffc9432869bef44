"""Bottleneck experiment: time selected conv layers at 1080p under FILM_DBG_FLAGS variants."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from frame_interpolation_b200 import synthetic
    from frame_interpolation_b200.interpolator import Interpolator
    eng = Interpolator("synthetic", align=64)
    x0, x1 = synthetic.frame_pair(1080, 1920, seed=0, n_waves=4)
    d0, d1 = torch.from_numpy(x0).cuda(), torch.from_numpy(x1).cuda()
    out = torch.empty_like(d0); torch.cuda.synchronize()
    eng.set_option("time_ops", 1)
    acc = None
    for i in range(3):
        eng.interpolate_device(d0.data_ptr(), d1.data_ptr(), 1, 1080, 1920, out.data_ptr()); eng.synchronize()
        t = {r["name"]: r["ms"] for r in eng.op_table()}
        acc = t if i == 0 else {k: min(acc[k], t[k]) for k in t}
    names = ["fe_conv1@L0", "flow_conv0@L0", "flow_conv1@L0", "fusion_conv1@L0", "fusion_conv2@L0", "flow_conv0@L1", "fe_conv3@L1", "flow_conv0@L3", "fusion_conv1@L2"]
    print("RES", os.environ.get("FILM_DBG_FLAGS", "0"), os.environ.get("FILM_DBG_NA", "-"), " ".join(f"{n}={acc[n]:.3f}" for n in names), "total=%.2f" % sum(acc.values()), flush=True)
else:
    for flags, na in [("0", None), ("1", None), ("2", None), ("4", None), ("3", None), ("7", None), ("0", "1")]:
        env = dict(os.environ, FILM_DBG_FLAGS=flags)
        if na: env["FILM_DBG_NA"] = na
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=300)
        print((r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip()[-300:] if r.returncode else "", flush=True)
