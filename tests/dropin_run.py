"""Run one small simulation through a libffengine-compatible library and dump the particles.
TEST INFRASTRUCTURE: python tests/dropin_run.py <lib.so> <out.npz> <flip|apic> [frames] [n] [features]
features: comma list of open (x+ side open), lifetime (fluid-particle lifetime attribute), surfvel (surface
velocity attribute against obstacles: the second VelocityAdvector call site, fluidsimulation.cpp:6951-6977), age /
viscosity / color (surface attributes: the AttributeToGridTransfer<float | vec3> call sites)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from blender_flip_fluids_b200 import scenes  # noqa: E402
from ffengine_mini import Engine  # noqa: E402

lib, out, method = sys.argv[1], sys.argv[2], sys.argv[3]
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n = int(sys.argv[5]) if len(sys.argv) > 5 else 24
features = set(sys.argv[6].split(",")) if len(sys.argv) > 6 and sys.argv[6] else set()
dx = 0.02 if n == 24 else 1.0 / n
sc = scenes.dam_break(n, apic=(method == "apic"), dx=dx, vel="swirl", v0=0.4, seed=17)
e = Engine(lib, n, n, n, sc.dx)
e.disable_console_output()
e.disable_surface_reconstruction()
if method == "apic":
    e.set_apic()
e.set_picflip_ratio(0.05)
e.set_max_thread_count(int(os.environ.get("FFB200_TEST_THREADS", "4")))
e.add_body_force(0.0, -9.81, 0.0)
if "open" in features:
    e.set_fluid_boundary_collisions([1, 0, 1, 1, 1, 1])
if "lifetime" in features:
    e.enable_fluid_particle_lifetime_attribute()
if "age" in features:            # AttributeToGridTransfer<float>, radius 1 dx (fluidsimulation.cpp:6991-7015)
    e.enable_surface_age_attribute()
if "viscosity" in features:      # AttributeToGridTransfer<float>, radius 3 dx (:7091-7115)
    e.enable_surface_viscosity_attribute()
if "color" in features:          # AttributeToGridTransfer<vmath::vec3> (:7148-7172)
    e.enable_surface_color_attribute()
if "surfvel" in features:
    e.enable_surface_velocity_attribute()
    e.enable_surface_velocity_attribute_against_obstacles()
e.load_marker_particle_data(sc.pos, sc.vel)
if method == "apic":
    e.load_marker_particle_affine_data(sc.affx * 0.01, sc.affy * 0.01, sc.affz * 0.01)
e.initialize()
wall = []
for _ in range(frames):
    t0 = time.perf_counter()
    e.update(1.0 / 60.0)
    wall.append(time.perf_counter() - t0)
st = e.frame_stats()
res = dict(pos=e.positions(), vel=e.velocities())
if method == "apic":
    res["affx"] = e.affinex()
np.savez(out, **res)
print("particles", e.num_marker_particles())
print("STATS " + json.dumps({"frame": st.frame, "substeps": st.substeps, "fluid_particles": st.fluid_particles,
                             "timing": {k: getattr(st.timing, k) for k, _ in st.timing._fields_}, "wall_s": wall}))
e.close()
