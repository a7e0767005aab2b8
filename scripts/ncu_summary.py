#!/usr/bin/env python
"""Turn `ncu` outputs brought back in gpurun_out/ into the committed text summaries under profiles/.

  python scripts/ncu_summary.py rep   <file.ncu-rep> <out.txt> [workload n world]   # --set full capture -> key metrics
  python scripts/ncu_summary.py list  <launches.csv> <out.txt>                      # launch list -> per-kernel totals

`rep` with a workload name also records dram bytes per launch of the LAST kernel in the report in profiles/traffic.json
(read by bench.py for roofline.traffic).
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
]

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    return head, units, data


def do_rep(rep, out_path, workload=None, n=None, world=1):
    head, units, data = raw_rows(rep)
    col = {}
    for i, name in enumerate(head):
        col.setdefault(name, i)
        col.setdefault(name.split(".", 2)[-1] if name.count(".") > 1 and name.split(".")[1].startswith("Triage") else name, i)
    lines = ["# summary of %s  (ncu --set full --clock-control none; per-launch, cold-cache, serialised)" % os.path.basename(rep)]
    last = None
    for row in data:
        lines.append("")
        lines.append("kernel: " + row[col["Kernel Name"]][:200])
        for key in KEYS:
            if key in col and row[col[key]] != "":
                lines.append("  %-92s %s %s" % (key, row[col[key]], units[col[key]]))
        rd = float(row[col["dram__bytes_read.sum"]]) * UNIT.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr = float(row[col["dram__bytes_write.sum"]]) * UNIT.get(units[col["dram__bytes_write.sum"]], 1.0)
        lines.append("  %-92s %.0f byte" % ("dram bytes (read+write) per launch", rd + wr))
        last = rd + wr
    open(out_path, "w").write("\n".join(lines) + "\n")
    if workload and last is not None:
        tpath = os.path.join(os.path.dirname(os.path.abspath(out_path)), "traffic.json")
        try:
            t = json.load(open(tpath))
        except (OSError, ValueError):
            t = {}
        t[workload] = {"dram_bytes_per_launch": last, "n": int(n), "world": int(world), "source": os.path.basename(out_path)}
        json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


def do_list(csv_path, out_path):
    rows = [r for r in csv.reader(l for l in open(csv_path) if not l.startswith("==")) if r]
    head = rows[0]
    ik, im, iv, iu = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(r[iu], 1.0)
        name = r[ik]
        short = name.split("(")[0][-110:]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    lines = ["# launch list summary of %s (ncu --metrics gpu__time_duration.sum --clock-control none; times are cold-cache," % os.path.basename(csv_path),
             "# serialised per-launch durations, so only the SHARE is comparable to the live CUDA-event numbers in bench.py)",
             "%-112s %6s %12s %7s" % ("kernel", "count", "total_us", "share")]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-112s %6d %12.1f %6.1f%%" % (k, c, v, 100.0 * v / max(total, 1e-9)))
    lines.append("%-112s %6s %12.1f" % ("TOTAL", "", total))
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        do_rep(*sys.argv[2:])
    else:
        do_list(*sys.argv[2:])
