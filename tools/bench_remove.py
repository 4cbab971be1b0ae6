"""Measurement of the widened row f2 (marker-particle removal), one JSON line.

    python tools/bench_remove.py [--n 128] [--steps 10] [--no-cpu]

Scene: the benchmark's dam break with random velocities (a few of them extreme) and a sphere obstacle dropped
into the fluid, so that all three rules of _removeMarkerParticles fire. GPU: ffb200_remove_marker_particles on the
resident particles, timed with CUDA events on the library's stream (the particles are re-uploaded, untimed, before
every repetition: the call is destructive); the time includes the call's one device->host read of the counts.
Roofline: algorithmic traffic = 3 passes over the velocities + 1 over the positions + reading and writing every
stream of the survivors once. CPU: the unmodified reference (oracle/_ref/ref_harness remove) on the same
particles -- only if the harness was built; its survivors are compared bit for bit with the GPU's.
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
    sc = scenes.dam_break(n, apic=True, vel="random", v0=0.5)
    dx, dt, cfl = sc.dx, 1.0 / 60.0, 5.0
    L = n * dx
    phi, near = scenes.analytic_solid_sdf(n, n, n, dx, sphere=(0.3 * L, 0.25 * L, 0.5 * L, 0.08 * L))
    rng = np.random.default_rng(3)
    vel = sc.vel.copy()
    vel[rng.choice(len(vel), 20, replace=False)] *= np.float32(200.0)
    stream = torch.cuda.current_stream()
    times = []
    with engine.FlipContext(n, n, n, dx) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_solid(phi, near)
        for it in range(a.steps + 3):
            ctx.set_particles(sc.pos, vel, sc.affx, sc.affy, sc.affz)
            ctx.sort_particles()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            remaining, extreme = ctx.remove_marker_particles(dt, cfl)
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        gp, gv, *_ = ctx.get_particles()
    total = len(vel)
    ms = float(np.median(times))
    alg = total * 12 * 4 + remaining * 15 * 4 * 2
    peak = 6539.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    line = {"metric": "marker-particle removal (_removeMarkerParticles)", "grid": [n, n, n], "particles": int(total),
            "remaining": int(remaining), "extreme_removed": int(extreme), "gpu_ms": ms, "particles_per_s": total / (ms * 1e-3),
            "roofline": {"bound": "hbm", "algorithmic_bytes": int(alg), "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": None}
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not a.no_cpu and os.path.exists(harness):
        d = tempfile.mkdtemp(prefix="ffremove_")
        np.save(os.path.join(d, "in_pos.npy"), sc.pos)
        np.save(os.path.join(d, "in_vel.npy"), vel)
        np.save(os.path.join(d, "in_phi.npy"), phi)
        r = subprocess.run([harness, "remove", d, f"I={n}", f"J={n}", f"K={n}", f"dx={dx!r}", f"dt={dt!r}", f"cfl={cfl}"],
                           capture_output=True, text=True)
        info = json.loads(r.stdout.strip().splitlines()[-1])
        rp, rv = np.load(os.path.join(d, "out_pos.npy")), np.load(os.path.join(d, "out_vel.npy"))
        same = rp.tobytes() == gp.tobytes() and rv.tobytes() == gv.tobytes() and info["extreme"] == extreme
        line["cpu_baseline"] = {"kind": "reference", "ms": info["t_remove"] * 1e3, "threads": info["threads"],
                                "bit_identical_to_gpu": bool(same)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
