"""profiles/<tag>_ncu_<name>.csv + a markdown table from the ncu reports a GPU call brought back in gpurun_out/.

    python tools/summarize_ncu_r2.py <tag> <report-stem> [<report-stem> ...]     e.g.  r2b pair single misc all

For every gpurun_out/<tag>_<stem>.ncu-rep: the raw page (`ncu -i ... --page raw --csv`) reduced to the columns the
roofline discussion uses, written as profiles/<tag>_ncu_<stem>.csv (one row per captured launch), and a short
markdown table appended to profiles/<tag>_ncu_summary.md."""
import csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, stems = sys.argv[1], sys.argv[2:]
WANT = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_bytes"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__cluster_size", "cluster"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct")]


def short(full):
    n = full.split("(const")[0].split("(float")[0].replace("void ", "")
    n = n.replace("film::<unnamed>::", "").replace("film::(anonymous namespace)::", "").replace("film::", "")
    return re.sub(r"\(.*", "", n)


def to_base(v, u):
    try:
        v = float(v.replace(",", ""))
    except ValueError:
        return v
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1,
            "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(u)
    return v * mult if mult else v


md = [f"# ncu evidence, build {tag}\n", "Captured with `tools/gpu_*.sh` (eager launches, `--clock-control none`); one row per launch, schedule order.",
      "Times under ncu are cold-cache and serialised: compare shares and counters, not absolutes.\n"]
for stem in stems:
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{stem}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if len(rr) < 3:
        print("empty", rep)
        continue
    hdr, units = rr[0], rr[1]
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    out_rows = []
    for r in rr[2:]:
        row = {}
        for i, n in cols:
            row[n] = short(r[i]) if n == "kernel" else to_base(r[i], units[i])
        out_rows.append(row)
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_{stem}.csv")
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=[n for _, n in cols])
        w.writeheader()
        w.writerows(out_rows)
    md.append(f"## `{tag}_{stem}.ncu-rep` ({len(out_rows)} launches) -> `profiles/{tag}_ncu_{stem}.csv`\n")
    md.append("| # | kernel | ms | tensor pipe % | L2 % | DRAM % | DRAM rd+wr GB | L2->SM GB | issue active % |")
    md.append("|---:|---|---:|---:|---:|---:|---:|---:|---:|")
    for i, r in enumerate(out_rows):
        g = lambda k, d=0.0: r.get(k, d) if isinstance(r.get(k, d), float) else d
        md.append(f"| {i} | `{r.get('kernel', '?')}` | {g('time') * 1e3:.3f} | {g('tensor_pipe_pct'):.1f} | {g('l2_pct'):.1f} | "
                  f"{g('dram_pct'):.1f} | {(g('dram_read') + g('dram_write')) / 1e9:.3f} | {g('l2_to_sm_bytes') / 1e9:.2f} | {g('issue_active_pct'):.1f} |")
    md.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md)[:3000])
