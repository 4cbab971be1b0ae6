"""Measurement of the widened row f3 (liquid SDF from particles), one JSON line.

    python tools/bench_liquid_sdf.py [--n 128] [--steps 10] [--no-cpu]

GPU: ffb200_liquid_sdf on the resident, sorted particles of the benchmark's dam break, timed with CUDA events on the
library's stream (the field stays on the device). Roofline: algorithmic traffic = the positions read once + the
cell grid written (fill), lowered and rewritten (decode); the scatter itself is bound by L2 read-modify-write
traffic, not HBM. CPU: the unmodified reference (oracle/_ref/ref_harness liquidsdf, all host threads) on the same
positions -- only if the harness was built; its field is compared bit for bit with the GPU's.
"""
import argparse, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from blender_flip_fluids_b200 import engine, scenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    n = a.n
    sc = scenes.dam_break(n, apic=False, vel="random", v0=0.5)
    dx = sc.dx
    radius = float(0.5 * dx * np.sqrt(3.0))
    stream = torch.cuda.current_stream()
    times = []
    with engine.FlipContext(n, n, n, dx) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_particles(sc.pos, sc.vel)
        ctx.sort_particles()
        for it in range(a.steps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.liquid_sdf(radius, download=False)
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        phi = ctx.liquid_sdf(radius)
    total = len(sc.pos)
    ms = float(np.median(times))
    alg = total * 12 + phi.size * 4 * 3
    peak = 6539.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    line = {"metric": "liquid SDF from particles (ParticleLevelSet::calculateSignedDistanceField)", "grid": [n, n, n],
            "particles": int(total), "cells_in_band": int((phi < np.float32(3.0 * dx)).sum()), "gpu_ms": ms,
            "particles_per_s": total / (ms * 1e-3),
            "roofline": {"bound": "hbm", "algorithmic_bytes": int(alg), "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": None}
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not a.no_cpu and os.path.exists(harness):
        d = tempfile.mkdtemp(prefix="ffsdf_")
        np.save(os.path.join(d, "in_pos.npy"), sc.pos)
        r = subprocess.run([harness, "liquidsdf", d, f"I={n}", f"J={n}", f"K={n}", f"dx={float(dx)!r}", f"radius={radius!r}", "reps=3"],
                           capture_output=True, text=True)
        info = json.loads(r.stdout.strip().splitlines()[-1])
        ref = np.load(os.path.join(d, "out_phi.npy"))
        line["cpu_baseline"] = {"kind": "reference", "ms": info["t_sdf"] * 1e3, "threads": info["threads"],
                                "bit_identical_to_gpu": bool(ref.tobytes() == phi.tobytes())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
