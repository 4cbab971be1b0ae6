"""Quick device-side stage timing (not the bench): python tools/time_stages.py N method [reps]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from blender_flip_fluids_b200 import engine, scenes

n = int(sys.argv[1]); method = sys.argv[2]; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
apic = method == "apic"
t0 = time.time()
sc = scenes.dam_break(n, apic=apic, vel="random", v0=0.5)
print(f"scene {sc.name}: {sc.n} particles ({time.time()-t0:.1f}s to build)", flush=True)
m = engine.APIC if apic else engine.FLIP
phi, near = scenes.analytic_solid_sdf(n, n, n, sc.dx)
dt = 1.0 * sc.dx / 0.5
with engine.FlipContext(n, n, n, sc.dx) as ctx:
    ctx.set_solid(phi, near)
    for r in range(reps):
        ctx.set_particles(sc.pos, sc.vel, sc.affx, sc.affy, sc.affz)
        ctx.p2g(sc.radius, m)
        ctx.save_velocity_field()
        ctx.g2p(m, 0.05)
        ctx.advect(dt, 5.0, True)
        t = ctx.timing()
        tot = t["sort_ms"] + t["p2g_prep_ms"] + t["p2g_ms"] + t["g2p_ms"] + t["advect_ms"]
        print(f"rep {r}: sort {t['sort_ms']:.3f}  prep {t['p2g_prep_ms']:.3f}  p2g {t['p2g_ms']:.3f}  g2p {t['g2p_ms']:.3f}  advect {t['advect_ms']:.3f}  "
              f"total {tot:.3f} ms -> {sc.n/tot/1e6:.2f} G particle-updates/s  (h2d {t['h2d_ms']:.2f} ms) launches "
              f"{t['sort_launches']}+{t['p2g_launches']}+{t['g2p_launches']}+{t['advect_launches']}", flush=True)
