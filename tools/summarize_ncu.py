"""Reduce ncu outputs to the markdown tables committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_bench_rX.csv          > table of the launch list
    python tools/summarize_ncu.py full     gpurun_out/prof_rX.ncu-rep                > one row per kernel of a --set full report
"""
import collections
import csv
import subprocess
import sys


def short(name):
    name = name.split("(")[0]
    for junk in ("ffb200::<unnamed>::", "void ", "unnamed>::", "ffb200::"):
        name = name.replace(junk, "")
    return name


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(short(r["Kernel Name"]), []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    return "\n".join(out)


COLS = [("time us", "gpu__time_duration.sum"), ("warp-instr", "smsp__inst_executed.sum"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("lanes/instr", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("fp64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("DRAM read MB", "dram__bytes_read.sum"), ("DRAM write MB", "dram__bytes_write.sum"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
        ("regs", "launch__registers_per_thread"), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("long-scoreboard stall", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3}
    out = ["| kernel | " + " | ".join(c for c, _ in COLS) + " |", "|---|" + "---|" * len(COLS)]
    seen = set()
    for r in rows[2:]:
        name = short(r[iname])
        if name in seen:
            continue
        seen.add(name)
        cells = []
        for label, metric in COLS:
            if metric not in hdr:
                cells.append("-")
                continue
            j = hdr.index(metric)
            v = float(r[j].replace(",", "")) * scale.get(units[j], 1)
            cells.append(f"{v / 1e6:.1f} M" if label == "warp-instr" else (f"{int(v)}" if label == "regs" else f"{v:.1f}"))
        out.append(f"| {name} | " + " | ".join(cells) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    print(launches(sys.argv[2]) if sys.argv[1] == "launches" else full(sys.argv[2]))
