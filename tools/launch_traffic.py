#!/usr/bin/env python
"""Launch list with DRAM bytes -> markdown table (stdout) and total DRAM bytes of the listed launches.

    python tools/launch_traffic.py profiles/r2_launches_512.csv [marker_kernel]

marker_kernel (default k_keys_rank, the first kernel of a step): only the launches from its first to just before its
last occurrence are used, i.e. whole steps.

Input: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv`.
Times under ncu are cold-cache and serialised: the shares are the evidence, not the absolutes."""
import collections
import csv
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
         "msecond": 1e3}


def short(name):
    name = name.split("(")[0]
    for junk in ("ffb200::<unnamed>::", "void ", "unnamed>::", "ffb200::"):
        name = name.replace(junk, "")
    return name


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "k_keys_rank"
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    per = collections.OrderedDict()           # launch id -> dict
    for r in rows:
        d = per.setdefault(r["ID"], {"name": short(r["Kernel Name"])})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"], 1.0)
    launches = list(per.values())
    marks = [i for i, d in enumerate(launches) if d["name"] == marker]
    steps = None
    if len(marks) >= 2:
        launches = launches[marks[0]:marks[-1]]
        steps = len(marks) - 1
    agg = collections.OrderedDict()
    for d in launches:
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print("| kernel | launches | mean us | share of listed time | DRAM read MB / launch | DRAM write MB / launch | GB/s |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        print(f"| `{k}` | {n} | {a['us'] / n:.1f} | {100 * a['us'] / tot:.1f} % | {a['rd'] / n / 1e6:.1f} | {a['wr'] / n / 1e6:.1f} | {gbs:.0f} |")
    total = sum(a["rd"] + a["wr"] for a in agg.values())
    print(f"\nlisted launches: {sum(a['n'] for a in agg.values())}, {tot / 1e3:.2f} ms, DRAM {total / 1e9:.2f} GB" +
          (f"; per step ({steps} whole steps): {tot / steps / 1e3:.2f} ms, {total / steps / 1e9:.2f} GB" if steps else ""))


if __name__ == "__main__":
    main()
