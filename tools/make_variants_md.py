"""profiles/<tag>_variants_ab.md from the same-box A/B bench runs of tools/gpu_full_r2.sh (gpurun_out/<tag>_bench_<cfg>.json)."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = []
for p in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_bench_*.json"))):
    cfg = os.path.basename(p)[len(tag) + 7:-5]
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        rows.append((cfg, None))
        continue
    rows.append((cfg, d))
base = dict(rows).get("default")
out = [f"# Kernel-option A/B on ONE box, build {tag} (1080p, 20 steps after 3 warm-ups, device-resident)\n",
       "Every line is `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-workloads` with one environment variable",
       "changed; all runs back to back on the same B200 (only same-box numbers are comparable: boxes differ by up to 8 % in",
       "sustained clock).\n",
       "| configuration | ms / step | frames/s | vs default | conv ms | roofline frac | gather GB/s | SM MHz |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
for cfg, d in rows:
    if d is None:
        out.append(f"| {cfg} | failed | | | | | | |")
        continue
    rel = d["ms_per_step"] / base["ms_per_step"] if base else float("nan")
    out.append(f"| {cfg} | {d['ms_per_step']:.3f} | {d['value']:.2f} | {rel:.3f}x | {d['roofline']['conv_kernel_ms_per_step']:.3f} | "
               f"{d['roofline']['frac']:.3f} | {d['gather']['achieved']:.0f} | {d['clocks'].get('sm_mhz')} |")
open(os.path.join(ROOT, "profiles", f"{tag}_variants_ab.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
