#!/usr/bin/env python
"""Summarise ncu captures brought back in gpurun_out/<tag>/ into profiles/<tag>_*.{md,csv} (tracked).

    python tools/ncu_summary.py r01a

Reads every *.ncu-rep of the tag with `ncu -i ... --page raw --csv` (works without a GPU), keeps the metrics the
roofline argument uses, and folds launches.csv (the `--metrics gpu__time_duration.sum` pass) into per-kernel shares.
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = OrderedDict()
        d["kernel"] = r[hdr.index("Kernel Name")]
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i], units[i])
        res.append(d)
    return res


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    lines = ["# ncu summary, capture `%s`" % tag, "",
             "Source: `ncu --set full --clock-control none --import-source on` under gpurun (1 x B200), read here with "
             "`ncu -i <rep> --page raw --csv` (tools/ncu_summary.py).  Times under ncu are serialised and cold-cache; "
             "bench.py's CUDA-event numbers are the ones of record.", ""]
    for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
        lines.append("## %s" % os.path.basename(rep))
        for d in raw_rows(rep):
            lines.append("")
            lines.append("**%s**" % d["kernel"])
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for k, v in d.items():
                if k != "kernel":
                    lines.append("| %s | %s | %s |" % (k, v[0], v[1]))
            try:
                rd = float(d["dram__bytes_read.sum"][0]); wr = float(d["dram__bytes_write.sum"][0])
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * scale[d["dram__bytes_read.sum"][1]] + wr * scale[d["dram__bytes_write.sum"][1]]
                t = float(d["gpu__time_duration.sum"][0]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[
                    d["gpu__time_duration.sum"][1]]
                lines.append("| **traffic (read+write)** | %.3f | MB |" % (tot / 1e6))
                lines.append("| **DRAM rate under ncu** | %.1f | GB/s |" % (tot / t / 1e9))
            except Exception:
                pass
        lines.append("")
    # launch list -> per-kernel shares
    lp = os.path.join(src, "launches.csv")
    if os.path.exists(lp):
        txt = open(lp).read()
        start = txt.find('"ID"')
        rows = list(csv.DictReader(io.StringIO(txt[start:]))) if start >= 0 else []
        agg = OrderedDict()
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            name = r["Kernel Name"].split("(")[0]
            a = agg.setdefault(name, [0.0, 0])
            a[0] += v
            a[1] += 1
        tot = sum(a[0] for a in agg.values()) or 1.0
        lines.append("## launch list (`--metrics gpu__time_duration.sum`): per-kernel totals")
        lines.append("")
        lines.append("| kernel | launches | total us | share |")
        lines.append("|---|---|---|---|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            lines.append("| %s | %d | %.1f | %.1f%% |" % (k, a[1], a[0], 100 * a[0] / tot))
        with open(os.path.join(dst, "%s_launches.csv" % tag), "w") as f:
            f.write(txt[start:] if start >= 0 else txt)
    bl = os.path.join(src, "bench_launches.csv")
    if os.path.exists(bl):   # the launch list of `python bench.py` itself
        txt = open(bl).read()
        start = txt.find('"ID"')
        rows = list(csv.DictReader(io.StringIO(txt[start:]))) if start >= 0 else []
        agg = OrderedDict()
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r.get("Metric Unit", "ns"), 1e-3)
            a = agg.setdefault(r["Kernel Name"].split("(")[0], [0.0, 0])
            a[0] += v
            a[1] += 1
        tot = sum(a[0] for a in agg.values()) or 1.0
        lines += ["", "## launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under ncu "
                  "(all legs: warm-up, timed steps, extras, end-to-end)", "", "| kernel | launches | total us | share |",
                  "|---|---|---|---|"]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            lines.append("| %s | %d | %.1f | %.1f%% |" % (k, a[1], a[0], 100 * a[0] / tot))
        with open(os.path.join(dst, "%s_bench_launches.csv" % tag), "w") as f:
            f.write(txt[start:] if start >= 0 else txt)
    bj = os.path.join(src, "bench.json")
    if os.path.exists(bj):
        try:
            b = json.load(open(bj))
            lines += ["", "## bench.py line of the same call (CUDA events, not under ncu)", "", "```json",
                      json.dumps(b, indent=1), "```"]
        except Exception:
            pass
    for extra in ("bench_ref.json", "gpu_tests.log", "smoke.log"):
        p = os.path.join(src, extra)
        if os.path.exists(p):
            lines += ["", "## %s" % extra, "", "```", open(p).read().strip()[-3000:], "```"]
    out = os.path.join(dst, "%s_ncu_summary.md" % tag)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main()
