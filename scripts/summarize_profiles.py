#!/usr/bin/env python3
"""Turn the ncu outputs of a gpurun pass into the small, tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag>      # e.g. r1c

reads  gpurun_out/launches.csv            (ncu --metrics gpu__time_duration.sum ... --csv)
       gpurun_out/prof_cg_fused_raw.csv   (ncu -i <report> --page raw --csv of the --set full capture)
writes profiles/<tag>_launches_raw.csv, profiles/<tag>_launch_list_summary.md,
       profiles/<tag>_ncu_full_summary.csv and refreshes profiles/spmv_dot_traffic.json.
"""
import csv
import json
import os
import re
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def csv_rows(path):
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    return list(csv.reader(lines))


def launch_list(tag, command):
    src = os.path.join(OUT, "launches.csv")
    if not os.path.exists(src):
        return
    raw = open(src).read()
    open(os.path.join(PROF, tag + "_launches_raw.csv"), "w").write(raw)
    rows = csv_rows(src)
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        if len(r) <= v:
            continue
        per.setdefault(short(r[k]), []).append(float(r[v].replace(",", "")) / 1e3)
    total = sum(sum(t) for t in per.values())
    lines = ["# %s launch list (ncu --metrics gpu__time_duration.sum --clock-control none, `%s`)" % (tag, command), "",
             "Per-launch times are cold-cache and serialised; compare SHARES.  Launches issued after the "
             "device latched `done` are ~7 us no-ops and pull the averages down; the medians are the working "
             "launches.  CUDA-graph replays appear as ordinary kernel launches.", "",
             "| kernel | launches | total us | share | median us |", "|---|---|---|---|---|"]
    for name, t in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append("| `%s` | %d | %.1f | %.1f %% | %.1f |" % (name[:110], len(t), sum(t), 100 * sum(t) / total,
                                                              statistics.median(t)))
    cg = {n: sum(t) for n, t in per.items() if "Cg" in n and "Setup" not in n and "Settle" not in n}
    spmv = sum(t for n, t in cg.items() if n.startswith("spmv_row_kernel"))
    if cg:
        lines += ["", "Share of the fused SpMV launch inside the CG step (the CG loop kernels only): %.1f %%."
                  % (100 * spmv / sum(cg.values()))]
    open(os.path.join(PROF, tag + "_launch_list_summary.md"), "w").write("\n".join(lines) + "\n")


def full_summary(tag, command):
    src = os.path.join(OUT, "prof_cg_fused_raw.csv")
    if not os.path.exists(src) or os.path.getsize(src) == 0:
        return
    rows = csv_rows(src)
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    names = [hdr[c] for c in cols]
    kcol = hdr.index("Kernel Name")
    out = ["# %s ncu --set full summaries (B200, BASELINE config 2: 5-pt Laplacian 3162^2, CG loop)" % tag,
           "# command: %s" % command,
           "# units: " + ", ".join("%s [%s]" % (n, units[c]) for n, c in zip(names, cols)),
           "kernel,launch," + ",".join(names)]
    seen = {}
    traffic = None
    for r in rows[2:]:
        if len(r) <= kcol:
            continue
        name = short(r[kcol])
        i = seen.get(name, 0)
        seen[name] = i + 1
        vals = [r[c].replace(",", "") for c in cols]
        out.append('"%s",%d,%s' % (name, i, ",".join(vals)))
        if name.startswith("spmv_row_kernel") and traffic is None:
            d = dict(zip(names, zip(vals, [units[c] for c in cols])))

            def to_bytes(key):
                val, unit = d[key]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                return float(val) * scale
            rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
            dur, dunit = d["gpu__time_duration.sum"]
            dur_us = float(dur) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(dunit, 1.0)
            traffic = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                       "duration_us": dur_us, "kernel": name, "cg_fuse": 2 if re.search(r"CgEpiFused<(1|true), (1|true)>", name)
                       else (1 if "CgEpiFused" in name else 0),
                       "source": "profiles/%s_ncu_full_summary.csv (ncu --set full, one launch, config 2)" % tag}
    open(os.path.join(PROF, tag + "_ncu_full_summary.csv"), "w").write("\n".join(out) + "\n")
    if traffic:
        n, nnz = 3162 * 3162, 5 * 3162 * 3162 - 4 * 3162
        traffic["algorithmic_bytes"] = 12 * nnz + 4 * (n + 1) + 16 * n + {0: 0, 1: 16, 2: 32}[traffic["cg_fuse"]] * n
        json.dump(traffic, open(os.path.join(PROF, "spmv_dot_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1c"
    os.makedirs(PROF, exist_ok=True)
    launch_list(tag, "python bench.py --steps 20 --warmup 3 --no-cpu")
    full_summary(tag, "ncu --set full --clock-control none --import-source on -k regex:'spmv_row_kernel|vec_pass_kernel' "
                      "-s 8 -c 4 python bench.py --steps 12 --warmup 3 --no-cpu")
    for f in ("bench_n1.json", "configs.jsonl", "ab_cgfuse.json", "clocks.csv"):
        src = os.path.join(OUT, f)
        if os.path.exists(src) and os.path.getsize(src):
            open(os.path.join(PROF, "%s_%s" % (tag, f)), "w").write(open(src).read())
