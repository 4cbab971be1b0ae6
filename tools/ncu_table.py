#!/usr/bin/env python
"""One row per kernel launch of an `ncu --set full` report (first `per_kernel` launches of each kernel).
    python tools/ncu_table.py gpurun_out/prof.ncu-rep [per_kernel]"""
import csv, subprocess, sys
path = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [("time us", "gpu__time_duration.sum"), ("warp-instr M", "smsp__inst_executed.sum"),
     ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("lanes/instr", "smsp__thread_inst_executed_per_inst_executed.ratio"),
     ("fp64 %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
     ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
     ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
     ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("l1tex %", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
     ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
     ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
     ("L2 hit %", "lts__t_sector_hit_rate.pct"), ("regs", "launch__registers_per_thread"),
     ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("long-sb stall", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     ("short-sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
     ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
     ("math-throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
     ("lg-throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
     ("not-selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
     ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")]
print("| kernel | " + " | ".join(m for m, _ in M) + " |")
print("|---|" + "---|" * len(M))
seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    for junk in ("ffb200::<unnamed>::", "void ", "unnamed>::", "ffb200::"):
        name = name.replace(junk, "")
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > per:
        continue
    out = [name]
    for m, k in M:
        if k not in idx:
            out.append("-"); continue
        v, u = r[idx[k]], units[idx[k]]
        try:
            f = float(v.replace(",", ""))
            if m == "time us":
                f *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1.0)
            if m == "warp-instr M": f /= 1e6
            if m.startswith("DRAM") and m.endswith("MB"):
                f = {"byte": f / 1e6, "Kbyte": f / 1e3, "Mbyte": f, "Gbyte": f * 1e3}.get(u, f)
            out.append(f"{f:.1f}")
        except ValueError:
            out.append(v)
    print("| " + " | ".join(out) + " |")
