"""Run one small simulation through a libffengine-compatible library and dump the particles.
TEST INFRASTRUCTURE: python tests/dropin_run.py <lib.so> <out.npz> <flip|apic> [frames]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from blender_flip_fluids_b200 import scenes  # noqa: E402
from ffengine_mini import Engine  # noqa: E402

lib, out, method = sys.argv[1], sys.argv[2], sys.argv[3]
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n = 24
sc = scenes.dam_break(n, apic=(method == "apic"), dx=0.02, vel="swirl", v0=0.4, seed=17)
e = Engine(lib, n, n, n, sc.dx)
e.disable_console_output()
e.disable_surface_reconstruction()
if method == "apic":
    e.set_apic()
e.set_picflip_ratio(0.05)
e.set_max_thread_count(4)
e.add_body_force(0.0, -9.81, 0.0)
e.load_marker_particle_data(sc.pos, sc.vel)
if method == "apic":
    e.load_marker_particle_affine_data(sc.affx * 0.01, sc.affy * 0.01, sc.affz * 0.01)
e.initialize()
for _ in range(frames):
    e.update(1.0 / 60.0)
res = dict(pos=e.positions(), vel=e.velocities())
if method == "apic":
    res["affx"] = e.affinex()
np.savez(out, **res)
print("particles", e.num_marker_particles())
e.close()
