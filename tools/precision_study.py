"""Error-budget study (CPU): emulate MMA operand formats per network stage in the oracle and
measure the end-to-end error against the fp64 oracle. Used to choose the tensor-core formulation
(DESIGN.md section 3)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frame_interpolation_b200 import weights, synthetic
from oracle.film_oracle import OracleInterpolator

def q(t, dt): return t.to(dt).to(torch.float32)
def split(x, dt):
    hi = q(x, dt); lo = q(x - hi, dt); return hi + lo
def make_hook(policy):
    # policy: dict stage -> mode ; stage in {feat, flow, fusion}; mode in {'f32','s16' (fp16 split),'b16' (bf16 split),'h1' (fp16 single)}
    def hook(x, k, name):
        st = 'feat' if name.startswith('feat_net') else 'flow' if name.startswith('predict_flow') else 'fusion'
        m = policy.get(st, 'f32')
        if m == 'f32': return x, k
        if m == 's16': return split(x, torch.float16), split(k, torch.float16)
        if m == 'b16': return split(x, torch.bfloat16), split(k, torch.bfloat16)
        if m == 'h1':  return q(split(x, torch.float16), torch.float16), q(k, torch.float16)   # activations stored split, hi plane only used
        raise ValueError(m)
    return hook

w = weights.synthetic_weights()
dt = np.full((1,), 0.5, np.float32)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for seed in (0, 1):
    x0, x1 = synthetic.frame_pair(size, size, seed)
    y64 = OracleInterpolator(w, align=64, dtype=torch.float64).interpolate(x0, x1, dt)
    for name, pol in [('all b16-split', dict(feat='b16', flow='b16', fusion='b16')),
                      ('all fp16-split', dict(feat='s16', flow='s16', fusion='s16')),
                      ('fusion fp16x1, rest fp16-split', dict(feat='s16', flow='s16', fusion='h1')),
                      ('feat+fusion fp16x1, flow split', dict(feat='h1', flow='s16', fusion='h1')),
                      ('flow fp16x1, rest split', dict(feat='s16', flow='h1', fusion='s16')),
                      ('all fp16x1', dict(feat='h1', flow='h1', fusion='h1'))]:
        y = OracleInterpolator(w, align=64, conv_hook=make_hook(pol)).interpolate(x0, x1, dt)
        e = np.abs(y.astype(np.float64) - y64)
        print(f'seed {seed} {name:34s} max-abs {e.max():.3e} mean {e.mean():.3e} p99.9 {np.quantile(e, 0.999):.3e}', flush=True)
