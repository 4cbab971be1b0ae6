"""Transfer-only sweep (BASELINE.json config #5) and per-config stage timings.

    python tools/sweep.py [--quick]

Prints one markdown table row per (grid, ppc, method): particles, device ms per stage and the
algorithmic GB/s per kernel group (SURVEY 8d bytes), measured with the library's CUDA events on
a fixed resident batch (3 warm-up + 5 timed substeps).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from blender_flip_fluids_b200 import engine, scenes

quick = "--quick" in sys.argv
cases = [(128, 8, "apic", "dam"), (128, 8, "flip", "dam"), (64, 8, "flip", "dam"),
         (256, 4, "flip", "dam"), (256, 4, "apic", "dam"), (256, 8, "flip", "dam"), (256, 8, "apic", "dam"),
         (256, 8, "flip", "fill")]
if not quick:
    cases += [(256, 27, "flip", "dam"), (256, 27, "apic", "dam")]
only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--method=")]
if only:
    cases = [c for c in cases if c[2] in only]


def balg(method, ppc):
    if method == "apic":
        return {"p2g": 60 + 15 / ppc, "g2p": 60 + 12 / ppc, "advect": 24 + 16 / ppc}
    return {"p2g": 24 + 15 / ppc, "g2p": 36 + 24 / ppc, "advect": 24 + 16 / ppc}


rows = []
print("| scene | grid | ppc | method | particles | sort ms | prep ms | p2g ms | g2p ms | advect ms | total ms | G upd/s | p2g GB/s | g2p GB/s | advect GB/s | step GB/s (alg) | % of 6540 |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, ppc, method, kind in cases:
    apic = method == "apic"
    t0 = time.time()
    sc = (scenes.fill_box if kind == "fill" else scenes.dam_break)(n, ppc=ppc, apic=apic, vel="random", v0=0.5, seed=1234)
    sphere = (0.5, 0.25, 0.5, 0.15) if kind == "fill" else None
    if sphere:
        d = np.linalg.norm(sc.pos.astype(np.float64) - np.array(sphere[:3]), axis=1)
        keep = d > sphere[3] + 0.5 * sc.dx
        sc.pos, sc.vel = sc.pos[keep], sc.vel[keep]
    phi, near = scenes.analytic_solid_sdf(n, n, n, sc.dx, sphere=sphere)
    m = engine.APIC if apic else engine.FLIP
    dt = 1.0 * sc.dx / 0.5
    with engine.FlipContext(n, n, n, sc.dx) as ctx:
        ctx.set_solid(phi, near)
        ctx.set_particles(sc.pos, sc.vel, sc.affx, sc.affy, sc.affz)
        ctx.set_fixed_batch(True)
        acc = {}
        for r in range(8):
            ctx.p2g(sc.radius, m)
            ctx.save_velocity_field()
            ctx.g2p(m, 0.05 if kind == "dam" else 0.02)
            ctx.advect(dt, 5.0, True)
            t = ctx.timing()
            if r >= 3:
                for k in ("sort_ms", "p2g_prep_ms", "p2g_ms", "g2p_ms", "advect_ms"):
                    acc[k] = acc.get(k, 0.0) + t[k] / 5.0
    tot = sum(acc.values())
    b = balg(method, ppc)
    N = sc.n
    gbs = lambda key, ms: b[key] * N / (ms * 1e-3) / 1e9
    step = sum(b.values()) * N / (tot * 1e-3) / 1e9
    row = dict(scene=kind, grid=n, ppc=ppc, method=method, particles=N, **acc, total_ms=tot, gups=N / tot / 1e6,
               p2g_gbs=gbs("p2g", acc["p2g_ms"]), g2p_gbs=gbs("g2p", acc["g2p_ms"]), advect_gbs=gbs("advect", acc["advect_ms"]),
               step_gbs=step)
    rows.append(row)
    print(f"| {kind} | {n}^3 | {ppc} | {method} | {N} | {acc['sort_ms']:.3f} | {acc['p2g_prep_ms']:.3f} | {acc['p2g_ms']:.3f} | "
          f"{acc['g2p_ms']:.3f} | {acc['advect_ms']:.3f} | {tot:.3f} | {N / tot / 1e6:.2f} | {row['p2g_gbs']:.0f} | "
          f"{row['g2p_gbs']:.0f} | {row['advect_gbs']:.0f} | {step:.0f} | {100 * step / 6539.9:.1f} |", flush=True)
    del sc
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/sweep.json", "w") as f:
    json.dump(rows, f, indent=1)
