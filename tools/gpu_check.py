"""Dev tool (runs on the GPU box): stage-by-stage comparison of the engine against the CPU oracle."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frame_interpolation_b200 import weights, synthetic, spec
from frame_interpolation_b200.interpolator import Interpolator
from oracle.film_oracle import OracleInterpolator

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="tc")
ap.add_argument("--h", type=int, default=128)
ap.add_argument("--w", type=int, default=128)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()

wpath = weights.ensure_synthetic_file()
w = weights.load(wpath)
x0, x1 = synthetic.frame_pair(a.h, a.w, a.seed)
dt = np.full((1,), 0.5, np.float32)
t = time.time()
aux = {}
ref = OracleInterpolator(w, align=64).interpolate(x0, x1, dt, aux)
print(f"oracle {time.time()-t:.2f}s", flush=True)

eng = Interpolator(wpath, align=64)
print(eng.version, flush=True)
eng.set_option("conv_impl", 1 if a.impl == "simt" else 0)
eng.set_option("use_graph", a.graph)
eng.set_option("keep_debug", 1)
t = time.time()
out = eng.interpolate(x0, x1, dt)
print(f"engine first call {time.time()-t:.3f}s  profile {eng.profile()}", flush=True)

def nhwc(t):  # oracle NCHW (1,C,H,W) -> flat NHWC
    return t[0].permute(1, 2, 0).contiguous().numpy().reshape(-1)

def cmp(name, got, want):
    want = want.astype(np.float64); got = got.astype(np.float64)
    e = np.abs(got - want)
    print(f"{name:18s} n={got.size:9d} max-abs {e.max():.3e} mean-abs {e.mean():.3e} ref-absmax {np.abs(want).max():.3e}", flush=True)
    return e.max()

H, W = ref.shape[1:3]
for l in range(7):
    for k in range(2):
        cmp(f"feat{k}/{l}", eng.debug_read(f"feat{k}/{l}"), nhwc(aux["feature_pyramids"][k][l]))
for l in reversed(range(7)):
    cmp(f"res_fwd/{l}", eng.debug_read(f"res_fwd/{l}"), nhwc(aux["forward_residual_flow_pyramid"][l]))
    cmp(f"res_bwd/{l}", eng.debug_read(f"res_bwd/{l}"), nhwc(aux["backward_residual_flow_pyramid"][l]))
for l in range(5):
    cmp(f"flow_fwd/{l}", eng.debug_read(f"flow_fwd/{l}"), nhwc(aux["forward_flow_pyramid"][l]))
    cmp(f"flow_bwd/{l}", eng.debug_read(f"flow_bwd/{l}"), nhwc(aux["backward_flow_pyramid"][l]))
    al = aux["aligned_pyramid"][l]
    C = spec.feature_channels(l)
    cmp(f"warped0/{l}", eng.debug_read(f"warped0/{l}"), nhwc(al[:, 3:3 + C]))
    cmp(f"warped1/{l}", eng.debug_read(f"warped1/{l}"), nhwc(al[:, 6 + C:6 + 2 * C]))
    side = torch.cat([al[:, 0:3], al[:, 3 + C:6 + C], al[:, 6 + 2 * C:10 + 2 * C]], dim=1)
    cmp(f"aligned_side/{l}", eng.debug_read(f"aligned_side/{l}"), nhwc(side))
m = cmp("image", out.reshape(-1), ref.reshape(-1))
psnr = lambda a, b: 10 * np.log10(1.0 / np.mean((a.astype(np.float64) - b) ** 2))
print(f"RESULT impl={a.impl} {a.h}x{a.w} max-abs={m:.3e} PSNR(engine vs oracle)={psnr(out, ref):.2f} dB")
for _ in range(3):
    out2 = eng.interpolate(x0, x1, dt)
print("repeat-bitwise", bool(np.array_equal(out, out2)), "profile", eng.profile())
