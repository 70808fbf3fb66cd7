#!/usr/bin/env python3
"""Summarise gpurun_out/*.ncu-rep + launches.csv into profiles/<tag>_*.{md,csv} (tracked).

usage: summarize_ncu.py <tag>      e.g. r01a
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def hot_lines(rep, top=18):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot_lines.py"), rep, str(top)], capture_output=True, text=True)
    return r.stdout


os.makedirs(PROF, exist_ok=True)
traffic = {}
launch_info = ("cfg2", 0)
try:
    w_, q_ = open(os.path.join(OUT, "profile_launch.txt")).read().split()
    launch_info = (w_, int(q_))
except (OSError, ValueError):
    pass
md = [f"# ncu summary {tag}", "",
      "Source: `tools/profile.sh` under gpurun (1x B200, `--clock-control none`). Full `.ncu-rep` files stay in",
      "`gpurun_out/` (scratch); this file holds what the numbers in DESIGN.md / bench.py are read from.", ""]
for name in ("bloom", "exact", "pairfilter", "dp", "rank", "score", "confusable"):
    rep = os.path.join(OUT, f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    # gpurun_out/ keeps the reports of earlier calls: only those written by THIS profile.sh run belong to the summary
    stamp = os.path.join(OUT, "profile_launch.txt")
    if os.path.exists(stamp) and os.path.getmtime(rep) + 1 < os.path.getmtime(stamp):
        continue
    m = raw_metrics(rep)
    try:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        traffic[f"{name}_kernel"] = {"dram_bytes": float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]],
                                     "workload": launch_info[0], "queries": launch_info[1],
                                     "source": f"profiles/{tag}_ncu_summary.md (ncu --set full, one launch)"}
        for key, metric in (("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            ("lanes_per_instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
                            ("warp_instructions", "smsp__inst_executed.sum"), ("lts_sectors_read", "lts__t_sectors_op_read.sum"),
                            ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")):
            if metric in m:
                try:
                    traffic[f"{name}_kernel"][key] = float(m[metric][0].replace(",", ""))
                except ValueError:
                    pass
    except (KeyError, ValueError):
        pass
    md.append(f"## {name}_kernel (`ncu --set full`, one launch of {launch_info[1]} {launch_info[0]} queries)")
    md.append("")
    md.append("| metric | value | unit |")
    md.append("|---|---|---|")
    for k in WANT:
        if k in m:
            md.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    md.append("")
    md.append("Hottest source lines (stall samples / warp instructions / avg active lanes):")
    md.append("")
    md.append("```")
    md.append(hot_lines(rep).rstrip())
    md.append("```")
    md.append("")
launch = os.path.join(OUT, "launches.csv")
if os.path.exists(launch):
    dst = os.path.join(PROF, f"{tag}_launches.csv")
    rows = [ln for ln in open(launch) if ln.startswith('"')]
    with open(dst, "w") as f:
        f.writelines(rows)
    tot = {}
    for r in csv.DictReader(io.StringIO("".join(rows))):
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        tot.setdefault(k, []).append(float(r["Metric Value"]))
    md.append("## launch list (gpu__time_duration.sum per launch, ns; cold-cache, serialised)")
    md.append("")
    md.append("| kernel | launches | median ns | share of the captured time |")
    md.append("|---|---|---|---|")
    import statistics
    med = {k: statistics.median(v) for k, v in tot.items()}
    s = sum(sum(v) for k, v in tot.items() if k != "encode_kernel")  # (encode runs once per batch, not per pass)
    for k, v in tot.items():
        share = "-" if k == "encode_kernel" else f"{sum(v) / s:.3f}"
        md.append(f"| {k} | {len(v)} | {med[k]:.0f} | {share} |")
    md.append("")
open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w").write("\n".join(md))
if "bloom_kernel" in traffic and "exact_kernel" in traffic:
    # candidate generation = Bloom stage + exact stage (bench.py's `roofline.traffic`)
    traffic["probe"] = dict(traffic["bloom_kernel"], dram_bytes=traffic["bloom_kernel"]["dram_bytes"] + traffic["exact_kernel"]["dram_bytes"])
if traffic and launch_info[1]:
    import json
    # one entry per workload: profiles/traffic.json = {workload: {kernel: {...}}}
    path = os.path.join(PROF, "traffic.json")
    try:
        allw = json.load(open(path))
        if "probe" in allw:  # round-1 layout (a single workload at the top level)
            allw = {allw["probe"].get("workload", "cfg2"): allw}
    except (OSError, ValueError):
        allw = {}
    key = launch_info[0] if launch_info[1] == 1_000_000 else f"{launch_info[0]}@{launch_info[1]}"  # (bench.py reads the 1 M-query entries)
    allw.setdefault(key, {}).update(traffic)  # (a partial capture only replaces the kernels it holds)
    json.dump(allw, open(path, "w"), indent=1)
print("wrote", os.path.join(PROF, f"{tag}_ncu_summary.md"))
