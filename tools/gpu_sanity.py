"""Sanity matrix on the GPU: every conv kernel variant x precision plan against the CPU oracle (small frames)."""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frame_interpolation_b200 import synthetic, weights
from frame_interpolation_b200.interpolator import Interpolator
from oracle.film_oracle import OracleInterpolator
import torch
torch.set_num_threads(16)
wpath = weights.ensure_synthetic_file()
orc = OracleInterpolator(weights.load(wpath), align=64)
dt = np.full((1,), 0.5, np.float32)
cases = [(256, 320, 13), (120, 180, 3), (576, 1024, 4)]   # the last one is large enough for the dual-item pair kernel
refs = {}
for h, w, s in cases:
    x0, x1 = synthetic.frame_pair(h, w, seed=s, n_waves=8)
    refs[(h, w, s)] = (x0, x1, orc(x0, x1, dt))
eng0 = Interpolator(wpath, align=64)
nst = len(eng0.stage_names())
dflt = eng0.get_option("onepass_default")
eng0.close()
ALL = (1 << nst) - 1
variants = [{}, {"conv3x3_2cta": 0}, {"conv3x3_2cta": 2}, {"conv3x3_v2": 0}, {"conv3x3_halo": 0}, {"conv3x3_halo": 1},
            {"conv3x3_halo": 3}, {"conv3x3_halo": 3, "conv3x3_2cta": 2}, {"fe_conv0_tc": 1}, {"fuse_rgb_head": 0},
            {"fe_conv0_tc": 1, "fuse_rgb_head": 0, "conv3x3_halo": 2}, {"conv3x3_dual": 1},
            {"conv3x3_dual": 1, "conv3x3_2cta": 2}, {"mma_straight": 0}, {"plane_skip": 0}, {"arena_reuse": 0}]
bad = 0
for var in variants:
    for mask in (0, dflt, ALL):
        eng = Interpolator(wpath, align=64)
        for k, v in var.items():
            eng.set_option(k, v)
        eng.set_option("onepass_mask", mask)
        errs = []
        for key, (x0, x1, ref) in refs.items():
            try:
                out = eng(x0, x1, dt)
                errs.append(float(np.abs(out.astype(np.float64) - ref).max()))
            except Exception as e:
                errs.append(float("nan"))
                print("  ERROR", var, hex(mask), key, e)
        lim = 2e-4 if mask == 0 else 2e-3
        ok = all(e < lim for e in errs)
        bad += 0 if ok else 1
        print(f"{'OK ' if ok else 'BAD'} {str(var):48s} mask {mask:#010x} max-abs " + " ".join(f"{e:.2e}" for e in errs), flush=True)
        eng.close()
print("bad:", bad)
sys.exit(1 if bad else 0)
