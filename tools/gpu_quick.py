"""One process = one engine configuration (environment variables FILM_*): small frames against the CPU oracle.
Prints `QUICK <tag> ok|FAIL ...`; exit code 0 iff every case is within tolerance. A wedged kernel traps after ~2 s
(bounded mbarrier waits) and only takes this process down."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
sizes = [(128, 192, 3), (256, 320, 13)] + ([(576, 1024, 4)] if os.environ.get("QUICK_BIG") else [])
from frame_interpolation_b200 import synthetic, weights
from frame_interpolation_b200.interpolator import Interpolator
from oracle.film_oracle import OracleInterpolator
import torch
torch.set_num_threads(16)
wpath = weights.ensure_synthetic_file()
orc = OracleInterpolator(weights.load(wpath), align=64)
dt = np.full((1,), 0.5, np.float32)
ok = True
res = []
try:
    for h, w, s in sizes:
        x0, x1 = synthetic.frame_pair(h, w, seed=s, n_waves=8)
        ref = orc(x0, x1, dt)
        for mask in (None, 0):
            eng = Interpolator(wpath, align=64)
            for kv in os.environ.get("QUICK_OPTS", "").split(","):
                if "=" in kv:
                    k, v = kv.split("=")
                    eng.set_option(k, int(v))
            if mask is not None:
                eng.set_option("onepass_mask", mask)
            out = eng(x0, x1, dt)
            e = float(np.abs(out.astype(np.float64) - ref).max())
            lim = 1e-4 if mask == 0 else 4e-4
            ok = ok and (e < lim)
            res.append(f"{h}x{w}/{'plan' if mask is None else 'mask0'}={e:.2e}")
            eng.close()
except Exception as exc:
    ok = False
    res.append(f"EXC {type(exc).__name__}: {str(exc)[:100]}")
print(f"QUICK {tag} {'ok' if ok else 'FAIL'} " + " ".join(res), flush=True)
sys.exit(0 if ok else 1)
