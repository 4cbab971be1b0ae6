"""Measurement of the widened row f1 (valid-face extrapolation + saved-field copy), one JSON line.

    python tools/bench_extrapolate.py [--n 128] [--steps 20] [--no-cpu]

GPU: ffb200_p2g leaves field + valid masks on the device; ffb200_extrapolate_velocity_field (12 layers) and
ffb200_save_velocity_field are timed with CUDA events on the library's stream. Roofline: the algorithmic
traffic of one layer is 1 status byte read + 1 written per face (+ 4 B written for the faces it fills),
against the measured HBM copy bandwidth. CPU: the unmodified reference (oracle/_ref/ref_harness extrapolate)
on the same field, all host threads -- only if the harness was built (it does not exist on a bare GPU box
unless oracle/_ref travelled with the snapshot).
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    n, layers = a.n, 12
    sc = scenes.dam_break(n, apic=True, vel="random", v0=0.5)
    stream = torch.cuda.current_stream()
    with engine.FlipContext(n, n, n, sc.dx) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_particles(sc.pos, sc.vel, sc.affx, sc.affy, sc.affz)
        ctx.set_fixed_batch(True)
        times = []
        for it in range(a.steps + 3):
            ctx.p2g(sc.radius, engine.APIC)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.extrapolate_velocity_field(layers)
            ctx.save_velocity_field()
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        ctx.p2g(sc.radius, engine.APIC)
        (u, v, w), (vu, vv, vw) = ctx.get_velocity_field()
    faces = u.size + v.size + w.size
    ms = float(np.median(times))
    alg = layers * faces * 2 + faces * 4 + faces * 8            # status bytes per layer + filled values + the saved copy
    peak = 6539.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    line = {"metric": "valid-face extrapolation + saved-field copy (GridUtils::extrapolateGrid x3, 12 layers)", "grid": [n, n, n],
            "faces": int(faces), "valid_faces": int(vu.sum() + vv.sum() + vw.sum()), "layers": layers, "gpu_ms": ms,
            "faces_per_s": faces / (ms * 1e-3),
            "roofline": {"bound": "hbm", "algorithmic_bytes": int(alg), "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": None}
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not a.no_cpu and os.path.exists(harness):
        d = tempfile.mkdtemp(prefix="ffextrap_")
        for name, arr in (("u", u), ("v", v), ("w", w)):
            np.save(os.path.join(d, f"in_{name}.npy"), arr)
        for name, arr in (("validu", vu), ("validv", vv), ("validw", vw)):
            np.save(os.path.join(d, f"in_{name}.npy"), arr.astype(np.bool_))
        r = subprocess.run([harness, "extrapolate", d, f"I={n}", f"J={n}", f"K={n}", f"dx={sc.dx!r}", f"layers={layers}", "reps=3"],
                           capture_output=True, text=True)
        info = json.loads(r.stdout.strip().splitlines()[-1])
        ref = [np.load(os.path.join(d, f"out_{c}.npy")) for c in "uvw"]
        with engine.FlipContext(n, n, n, sc.dx) as ctx:         # parity at full size against the unmodified reference
            ctx.set_velocity_field(u, v, w)
            ctx.set_valid_velocities(vu, vv, vw)
            ctx.extrapolate_velocity_field(layers)
            got, _ = ctx.get_velocity_field()
        same = all(a_.tobytes() == b_.tobytes() for a_, b_ in zip(got, ref))
        line["cpu_baseline"] = {"kind": "reference", "ms": info["t_extrapolate"] * 1e3, "threads": info["threads"],
                                "bit_identical_to_gpu": bool(same)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
