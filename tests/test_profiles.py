"""CPU suite: the committed evidence under profiles/ is self-consistent -- the per-step DRAM traffic bench.py reports
(`roofline.traffic`, read from profiles/step_traffic.json) follows from the committed ncu launch list, and the committed
bench lines of the 512^3 scene carry the same particle / field checksums at every N."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _line(name):
    with open(os.path.join(PROF, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_step_traffic_follows_from_the_launch_list():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_traffic.py"), os.path.join(PROF, "r2_launches_512.csv")],
                       capture_output=True, text=True, check=True)
    tail = r.stdout.strip().splitlines()[-1]
    assert "2 whole steps" in tail
    per_step_gb = float(tail.split("ms,")[-1].split("GB")[0])
    with open(os.path.join(PROF, "step_traffic.json")) as f:
        t = json.load(f)
    # kernels + bin-table memset (4.4 GB) + saved-field copy and field overwrite (6.5 GB)
    assert abs(t["dram_bytes_per_step"] / 1e9 - (per_step_gb + 4.4 + 6.5)) < 0.5
    assert t["algorithmic_bytes_per_step"] == 149.375 * 330341088


def test_committed_bench_lines_agree_across_decompositions():
    lines = {n: _line(f"r2_bench_n{n}.json") for n in (1, 2, 4, 8)}
    for n, d in lines.items():
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["workload"] == lines[1]["config"]["workload"]
        assert d["checksum"]["particles"] == 330341088 and d["checksum"]["after_steps"] == lines[1]["checksum"]["after_steps"]
        assert d["checksum"]["particle_hash"] == lines[1]["checksum"]["particle_hash"], n
        assert d["checksum"]["p2g_field_hash"] == lines[1]["checksum"]["p2g_field_hash"], n
        assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "hbm"
    long1, long8 = _line("r2_bench_n1_65steps.json"), _line("r2_bench_n8_65steps.json")
    assert long1["checksum"]["after_steps"] == long8["checksum"]["after_steps"] == 65
    assert long1["checksum"]["particle_hash"] == long8["checksum"]["particle_hash"]
    assert long1["checksum"]["p2g_field_hash"] == long8["checksum"]["p2g_field_hash"]
    assert lines[1]["ms_per_step"] / lines[8]["ms_per_step"] > 6.0          # the strong-scaling figure DESIGN.md quotes
    ref = _line("r2_bench_reference.json")
    assert ref["impl"] == "reference" and ref["config"]["workload"] == lines[1]["config"]["workload"]
    assert ref["cpu_baseline"]["same_config"] is True
