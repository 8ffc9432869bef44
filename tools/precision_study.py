"""Per-stage precision study of the engine's tensor-core convs, ON THE GPU, at full size.

    python tools/precision_study.py [--height 1080 --width 1920] [--seeds 0,1] [--out gpurun_out/precision]

Every conv call site belongs to a stage of the precision plan (film_stage_name). For each stage the study
switches ONLY that stage from the three-pass split product (hi*hi + hi*lo + lo*hi, fp32-grade) to the
single-pass product (hi*hi: fp16 operands, fp32 accumulate), and measures
  * the error of the final image against the fp32 CPU oracle (max-abs, rms, p99.9), and
  * the device time of one network call (CUDA events around `film_interpolate_device`, graph replay),
then adds stages greedily in the order of error-variance cost per millisecond saved and reports the
cumulative error / time curve, from which the default plan (`kDefaultOnepassMask`, film_engine.cu) is taken
with a >= 3x margin under the 1e-3 contract. The oracle is used here as the checker only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--seeds", default="0")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "precision_study"))
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--threads", type=int, default=32)
    ap.add_argument("--extra-masks", default="", help="comma-separated hex masks to evaluate as well")
    a = ap.parse_args()

    import torch
    from frame_interpolation_b200 import synthetic, weights
    from frame_interpolation_b200.interpolator import Interpolator
    from oracle.film_oracle import OracleInterpolator

    h, w = a.height, a.width
    seeds = [int(s) for s in a.seeds.split(",")]
    wpath = weights.ensure_synthetic_file()
    eng = Interpolator(wpath, align=64)
    names = eng.stage_names()
    default_mask = eng.get_option("onepass_default")
    dt = np.full((1,), 0.5, np.float32)
    torch.set_num_threads(min(a.threads, os.cpu_count() or 1))
    orc = OracleInterpolator(weights.load(wpath), align=64)

    pairs, refs = [], []
    for s in seeds:
        x0, x1 = synthetic.frame_pair(h, w, seed=s, n_waves=8)
        t = time.time()
        refs.append(orc(x0, x1, dt).astype(np.float64))
        print(f"oracle seed {s}: {time.time() - t:.1f} s", flush=True)
        pairs.append((torch.from_numpy(x0).cuda(), torch.from_numpy(x1).cuda()))
    dout = torch.empty_like(pairs[0][0])
    stream = torch.cuda.Stream()

    def measure(mask):
        eng.clear_cache()
        eng.set_option("onepass_mask", mask)
        errs = []
        for (d0, d1), ref in zip(pairs, refs):
            eng.interpolate_device(d0.data_ptr(), d1.data_ptr(), 1, h, w, dout.data_ptr(), stream=stream.cuda_stream)
            stream.synchronize()
            e = np.abs(dout.cpu().numpy().astype(np.float64) - ref)
            errs.append((float(e.max()), float(np.sqrt((e ** 2).mean())), float(np.quantile(e, 0.999))))
        d0, d1 = pairs[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.iters):
            eng.interpolate_device(d0.data_ptr(), d1.data_ptr(), 1, h, w, dout.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        return {"mask": mask, "max_abs": max(x[0] for x in errs), "rms": float(np.sqrt(np.mean([x[1] ** 2 for x in errs]))),
                "p999": max(x[2] for x in errs), "ms": ms}

    rows = []
    base = measure(0)
    base["name"] = "all three-pass"
    rows.append(base)
    print(json.dumps(base), flush=True)
    single = []
    for i, n in enumerate(names):
        r = measure(1 << i)
        r["name"] = n
        r["saved_ms"] = base["ms"] - r["ms"]
        r["var_cost"] = max(r["rms"] ** 2 - base["rms"] ** 2, 1e-16)
        single.append(r)
        rows.append(r)
        print(json.dumps(r), flush=True)
    # greedy cumulative curve: cheapest error variance per ms saved first; stages that save nothing go last
    order = sorted(range(len(names)), key=lambda i: single[i]["var_cost"] / max(single[i]["saved_ms"], 1e-3))
    cum, mask = [], 0
    for i in order:
        mask |= 1 << i
        r = measure(mask)
        r["name"] = "+ " + names[i]
        cum.append(r)
        print(json.dumps(r), flush=True)
    extra = []
    for m in [default_mask] + [int(x, 16) for x in a.extra_masks.split(",") if x]:
        r = measure(m)
        r["name"] = ("default plan" if m == default_mask else "mask") + f" {m:#x}: " + ",".join(
            n for i, n in enumerate(names) if (m >> i) & 1)
        extra.append(r)
        print(json.dumps(r), flush=True)

    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out + ".json", "w") as f:
        json.dump({"height": h, "width": w, "seeds": seeds, "version": eng.version, "stages": names,
                   "baseline": base, "single": single, "cumulative": cum, "extra": extra}, f, indent=1)
    with open(a.out + ".md", "w") as f:
        f.write(f"# Precision study {w}x{h}, seeds {seeds}, {eng.version}\n\n")
        f.write("Error of the final image vs the fp32 CPU oracle; ms = one network call (CUDA events, graph replay).\n\n")
        f.write("## One stage single-pass at a time\n\n| stage | max-abs | rms | p99.9 | ms | saved ms |\n|---|---:|---:|---:|---:|---:|\n")
        f.write(f"| (all three-pass) | {base['max_abs']:.2e} | {base['rms']:.2e} | {base['p999']:.2e} | {base['ms']:.3f} | |\n")
        for r in single:
            f.write(f"| {r['name']} | {r['max_abs']:.2e} | {r['rms']:.2e} | {r['p999']:.2e} | {r['ms']:.3f} | {r['saved_ms']:.3f} |\n")
        f.write("\n## Greedy cumulative (cheapest error variance per ms first)\n\n| added stage | mask | max-abs | rms | p99.9 | ms |\n|---|---|---:|---:|---:|---:|\n")
        for r in cum:
            f.write(f"| {r['name']} | {r['mask']:#x} | {r['max_abs']:.2e} | {r['rms']:.2e} | {r['p999']:.2e} | {r['ms']:.3f} |\n")
        f.write("\n## Named plans\n\n| plan | max-abs | rms | p99.9 | ms |\n|---|---:|---:|---:|---:|\n")
        for r in extra:
            f.write(f"| {r['name']} | {r['max_abs']:.2e} | {r['rms']:.2e} | {r['p999']:.2e} | {r['ms']:.3f} |\n")
    print("wrote", a.out + ".md")
    eng.close()


if __name__ == "__main__":
    main()
