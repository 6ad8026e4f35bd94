"""Turn gpurun_out ncu artefacts into the committed summaries under profiles/.

  python tools/profile_summary.py launches <launches.csv> <out.md>      per-kernel share of the step
  python tools/profile_summary.py kernels  <report.ncu-rep> <out.md> <out.json>   per-kernel counters
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("dpp::<unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[ki])
        t = float(r[vi].replace(",", "")) / 1e3  # ns -> us
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path.split('/')[-1]})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over a window of the timed solve of "
                "`python bench.py --steps 1 --warmup 3 --no-cpu` (256^3, Jacobi-CG; `-s 1500 -c 400`). Per-launch times are cold-cache and "
                "serialised: read the SHARE column.\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / total:.1f}% |\n")
        f.write(f"\nwindow total {total:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
    print(open(out).read())


KEYS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
])


def kernels(rep, out_md, out_json, what=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    summary = OrderedDict()
    with open(out_md, "w") as f:
        f.write(f"# ncu --set full capture ({rep.split('/')[-1]})\n\n")
        f.write("`ncu --set full --clock-control none --import-source on` on "
                + (what or "`python bench.py --steps 1 --warmup 1 --no-cpu` (256^3 hex Q1, Jacobi-CG)")
                + ". One row per captured launch.\n\n")
        cols = list(KEYS.values())
        f.write("| kernel | " + " | ".join(cols) + " | top stalls (per issue) |\n|---|" + "---:|" * len(cols) + "---|\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            vals = []
            rec = {}
            for k, lab in KEYS.items():
                if k in idx:
                    v, u = r[idx[k]], units[idx[k]]
                    rec[lab] = f"{v} {u}".strip()
                    try:
                        vals.append(f"{float(v.replace(',', '')):.4g} {u}".strip())
                    except ValueError:
                        vals.append(v)
                else:
                    vals.append("-")
            stalls = []
            for h, i in idx.items():
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "selected" not in h:
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            st = ", ".join(f"{n} {v:.2f}" for v, n in stalls[:4])
            f.write(f"| `{name}` | " + " | ".join(vals) + f" | {st} |\n")
            def num(key):
                return float(r[idx[key]].replace(",", "")) if key in idx else None
            def to_bytes(key):
                v = num(key)
                u = units[idx[key]].lower()
                mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                return v * mult
            summary.setdefault(name, []).append({
                "duration_us": num("gpu__time_duration.sum") / (1e3 if units[idx["gpu__time_duration.sum"]] == "ns" else 1),
                "dram_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
            })
    json.dump(summary, open(out_json, "w"), indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernels(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
