#!/usr/bin/env python
"""Summarise ncu output into small tracked files.

  python profiles/summarize.py launches <launches.csv> <out.md>     # per-kernel time shares from the launch list
  python profiles/summarize.py full <prof.ncu-rep> <out.md> [traffic.json]   # key metrics of a --set full capture
"""
import csv
import json
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("same::", "")
    name = re.sub(r"void cub::CUB_[0-9_A-Z]+::", "cub::", name)
    name = re.sub(r"cub::detail::[a-z_]+::", "cub::", name)
    return name[:90]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    NOT_STEP = ("k_fp64_fma", "FillFunctor", "k_resolve_", "k_bbox", "k_fill_f64", "k_iota")
    off_step = lambda k: any(t in k for t in NOT_STEP)
    tot = sum(v[1] for k, v in agg.items() if not off_step(k))
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised: compare SHARES.\n\n")
        f.write(f"{len(rows)} launches, {tot / 1e6:.3f} ms in step kernels (shares are of this)\n\n")
        f.write("Not part of a step: `k_fp64_fma` (FP64 peak micro-benchmark, after the timed region), `at::...FillFunctor` (the 512 MiB L2 "
                "flush between steps), `k_resolve_*` / `k_bbox` (section set-up).\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e3:.1f} | {ns / n / 1e3:.2f} | {'(not a step kernel)' if off_step(k) else f'{ns / tot:.3f}'} |\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def full(rep, out, traffic_path=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    units = rows[1]
    data = rows[2:]
    ki = hdr.index("Kernel Name")
    traffic, extra = {}, {}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n")
        for r in data:
            f.write(f"## `{short(r[ki])}`  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            vals = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    vals[k] = r[i]
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
            try:
                def tobytes(k):
                    i = hdr.index(k)
                    v = float(r[i].replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
                traffic.setdefault(short(r[ki]), []).append(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
            except Exception:
                pass
            for k, name in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                            ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_active_pct"),
                            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
                            ("lts__t_sector_hit_rate.pct", "l2_hit_pct")):
                if k in hdr:
                    try:
                        extra.setdefault(short(r[ki]).replace("void ", ""), {})[name] = float(r[hdr.index(k)].replace(",", ""))
                    except ValueError:
                        pass
    if traffic_path:
        old = {}
        try:
            old = json.load(open(traffic_path))
        except Exception:
            pass
        for k, v in traffic.items():
            old[k.replace("void ", "")] = sum(v) / len(v)
        json.dump(old, open(traffic_path, "w"), indent=1, sort_keys=True)
        mpath = traffic_path.replace("ncu_traffic", "ncu_metrics")
        oldm = {}
        try:
            oldm = json.load(open(mpath))
        except Exception:
            pass
        oldm.update(extra)
        json.dump(oldm, open(mpath, "w"), indent=1, sort_keys=True)
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
