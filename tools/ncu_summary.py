#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launch lists into the small tracked summaries under profiles/.

    python tools/ncu_summary.py shares  gpurun_out/r01b_launches.csv  profiles/r01b_launch_shares.csv  "<command>"
    python tools/ncu_summary.py kernel  gpurun_out/X.ncu-rep          profiles/X_ncu_summary.json     "<command>" "<workload>"

Runs here (no GPU needed): ``ncu -i`` only reads the report."""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "derived__memory_l1_wavefronts_shared_excessive",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "sm__cycles_active.avg", "smsp__inst_executed.sum",
]


def shares(src, dst, command):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.Counter(); cnt = collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
        tot[name] += int(r[iv].replace(",", "")); cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  {command}\n")
        fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        fh.write("kernel,launches,total_ns,share\n")
        for name, ns in tot.most_common():
            fh.write(f"{name},{cnt[name]},{ns},{ns / total:.4f}\n")
    print(open(dst).read())


def kernel(src, dst, command, workload):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            m[name] = {"value": vals[i], "unit": units[i]}
    name = vals[hdr.index("Kernel Name")]
    json.dump({"kernel": name, "command": command, "workload": workload, "metrics": m}, open(dst, "w"), indent=1)
    for k, v in m.items():
        print(f"{k:90s} {v['value']:>14s} {v['unit']}")


if __name__ == "__main__":
    {"shares": shares, "kernel": kernel}[sys.argv[1]](*sys.argv[2:])
