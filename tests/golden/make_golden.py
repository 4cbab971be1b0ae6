"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference engine.

Run in the build container only (needs /root/reference to build oracle/_ref):

    make -C oracle -j8 all && python tests/golden/make_golden.py

Every array in the fixtures is an output of oracle/_ref/ref_harness, i.e. of the reference's
own VelocityAdvector::advect, _updateMarkerParticleVelocitiesThread and
_advanceMarkerParticlesThread compiled from /root/reference/src/engine (v1.8.5). The reference
ships no tests or golden vectors of its own, so these are the pin for
oracle/flip_oracle.c and, through it, for the CUDA path. Inputs are seeded numpy arrays
(blender_flip_fluids_b200/scenes.py) so the script is reproducible.
"""
from __future__ import annotations

import json
import math
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blender_flip_fluids_b200 import scenes  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
OUT = os.path.dirname(os.path.abspath(__file__))


def run(mode, workdir, **kw):
    cmd = [HARNESS, mode, workdir] + [f"{k}={v!r}" if isinstance(v, float) else f"{k}={v}" for k, v in kw.items()]
    r = subprocess.run(cmd, capture_output=True, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


def save_inputs(d, **arrs):
    for k, a in arrs.items():
        if a is not None:
            np.save(os.path.join(d, f"in_{k}.npy"), a)


def load_all(d, prefix):
    out = {}
    for f in sorted(os.listdir(d)):
        if f.startswith(prefix) and f.endswith(".npy"):
            out[f[:-4]] = np.load(os.path.join(d, f))
    return out


def seeded_particles(I, J, K, dx, rng, apic):
    """Ragged blob with a free surface, spanning the 10-node block seams, 8-ish ppc."""
    sc = scenes.dam_break(max(I, J, K), apic=apic, seed=int(rng.integers(1 << 30)), dx=dx, vel="random", v0=1.0,
                          dims=(I, J, K))
    keep = rng.random(sc.n) < 0.85                       # knock holes in it: ragged fringe faces
    pos, vel = sc.pos[keep], sc.vel[keep]
    aff = [a[keep] for a in (sc.affx, sc.affy, sc.affz)] if apic else [None] * 3
    return pos, vel, aff


def seam_particles(I, J, K, dx, rng, count=600):
    """Adversarial positions: within a few ulps / 1e-6 of block seams (10*dx multiples, with
    and without the half-cell stagger) and of the 'simple vs overlapping' threshold
    seam -+ (radius + 1e-6) (velocityadvector.cpp:306-320), plus cell boundaries."""
    r = 0.5 * dx * math.sqrt(3.0)
    pts = []
    seams = [10 * dx * m for m in range(1, max(I, J, K) // 10 + 1)]
    for _ in range(count):
        p = np.array([rng.uniform(3, I - 3), rng.uniform(3, J - 3), rng.uniform(3, K - 3)]) * dx
        ax = int(rng.integers(3))
        s = seams[int(rng.integers(len(seams)))]
        if s >= (I, J, K)[ax] * dx - 3 * dx:
            s = seams[0]
        kind = int(rng.integers(6))
        off = [0.0, 0.5 * dx][int(rng.integers(2))]
        base = {0: s, 1: s - (r + 1e-6), 2: s + (r + 1e-6), 3: s - r, 4: s + r,
                5: dx * int(rng.integers(4, (I, J, K)[ax] - 4))}[kind] + off
        p[ax] = base + rng.choice([0.0, 1e-7, -1e-7, 3e-8, -3e-8, 1e-6, -1e-6, 1e-9]) * (1.0 if rng.random() < 0.5 else dx)
        pts.append(p)
    return np.asarray(pts, np.float32)


def fixture_scene(name, I, J, K, dx, method, warm, obstacle, dt, seed):
    rng = np.random.default_rng(seed)
    apic = method == "apic"
    pos, vel, aff = seeded_particles(I, J, K, dx, rng, apic)
    vel = (vel * 0.4).astype(np.float32)
    d = tempfile.mkdtemp(prefix="ffgold_")
    save_inputs(d, pos=pos, vel=vel, affx=aff[0], affy=aff[1], affz=aff[2])
    info = run("scene", d, I=I, J=J, K=K, dx=float(dx), method=method, warm=warm, obstacle=obstacle, dt=float(dt))
    arrs = load_all(d, "s")
    shutil.rmtree(d)
    meta = dict(I=I, J=J, K=K, dx=dx, method=method, dt=dt, ratio=info["ratio"], cfl=info["cfl"],
                radius=info["radius"], warm=warm, obstacle=obstacle, particles=info["particles"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrs)
    print(name, meta)


def fixture_p2g(name, I, J, K, dx, method, seed, radius_scale=1.0):
    rng = np.random.default_rng(seed)
    apic = method == "apic"
    pos, vel, aff = seeded_particles(I, J, K, dx, rng, apic)
    sp = seam_particles(I, J, K, dx, rng)
    pos = np.concatenate([pos, sp]).astype(np.float32)
    perm = rng.permutation(pos.shape[0])                 # seam particles interleaved in index order
    pos = pos[perm]
    vel = np.concatenate([vel, rng.uniform(-1, 1, size=sp.shape).astype(np.float32)])[perm]
    if apic:
        aff = [np.concatenate([a, (rng.uniform(-1, 1, size=sp.shape) * 0.1 / dx).astype(np.float32)])[perm] for a in aff]
    radius = 0.5 * dx * math.sqrt(3.0) * radius_scale
    d = tempfile.mkdtemp(prefix="ffgold_")
    save_inputs(d, pos=pos, vel=vel, affx=aff[0], affy=aff[1], affz=aff[2])
    run("p2g", d, I=I, J=J, K=K, dx=float(dx), method=method, radius=float(radius), threads=3)
    arrs = load_all(d, "out_")
    arrs.update(load_all(d, "in_"))
    shutil.rmtree(d)
    meta = dict(I=I, J=J, K=K, dx=dx, method=method, radius=radius, particles=int(pos.shape[0]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrs)
    print(name, meta)


def fixture_advect(name, scene_npz, seed, cells_per_step=4.0):
    """Collision-heavy advection: the reference's own solid SDF / near-solid grid from a scene
    fixture, and a coherent velocity field that drives particles into walls and obstacle."""
    z = np.load(os.path.join(OUT, scene_npz + ".npz"))
    meta = json.loads(str(z["meta"]))
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    rng = np.random.default_rng(seed)
    dt = 1.0 / 30.0
    vmag = cells_per_step * dx / dt
    shp = [(K, J, I + 1), (K, J + 1, I), (K + 1, J, I)]
    zz, yy, xx = np.meshgrid(np.arange(K + 1), np.arange(J + 1), np.arange(I + 1), indexing="ij")
    base = [np.sin(0.4 * yy + 0.3 * zz), -np.cos(0.35 * xx + 0.2 * zz) - 0.5, np.sin(0.5 * xx - 0.3 * yy)]
    mac = []
    for c, s in enumerate(shp):
        f = base[c][: s[0], : s[1], : s[2]] * vmag + rng.uniform(-0.3, 0.3, size=s) * vmag
        mac.append(f.astype(np.float32))
    pos = z["s0_pos"]
    d = tempfile.mkdtemp(prefix="ffgold_")
    save_inputs(d, pos=pos, vel=np.zeros_like(pos), u=mac[0], v=mac[1], w=mac[2], phi=z["s2_phi"], near=z["s2_near"])
    run("advect", d, I=I, J=J, K=K, dx=float(dx), dt=float(dt), cfl=5)
    out = np.load(os.path.join(d, "out_pos.npy"))
    shutil.rmtree(d)
    m2 = dict(I=I, J=J, K=K, dx=dx, dt=dt, cfl=5.0, particles=int(pos.shape[0]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(m2), in_pos=pos, in_u=mac[0], in_v=mac[1],
                        in_w=mac[2], in_phi=z["s2_phi"], in_near=z["s2_near"], out_pos=out)
    print(name, m2, "moved", int((out != pos).any(axis=1).sum()))


def fixture_extrapolate(name, scene_npz, layers, threads):
    """Valid-face extrapolation (GridUtils::extrapolateGrid via MACVelocityField::extrapolateVelocityField)
    of the reference's own P2G output + valid masks taken from a scene fixture. Run with several
    reference thread counts: the outputs must agree (the threaded passes are order-independent)."""
    z = np.load(os.path.join(OUT, scene_npz + ".npz"))
    meta = json.loads(str(z["meta"]))
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    outs = []
    for t in threads:
        d = tempfile.mkdtemp(prefix="ffgold_")
        save_inputs(d, u=z["s1_u"], v=z["s1_v"], w=z["s1_w"], validu=z["s1_validu"], validv=z["s1_validv"],
                    validw=z["s1_validw"])
        run("extrapolate", d, I=I, J=J, K=K, dx=float(dx), layers=layers, threads=t)
        outs.append(load_all(d, "out_"))
        shutil.rmtree(d)
    for o in outs[1:]:
        for k in outs[0]:
            assert o[k].tobytes() == outs[0][k].tobytes(), f"reference extrapolation depends on the thread count ({k})"
    m2 = dict(I=I, J=J, K=K, dx=dx, layers=layers, scene=scene_npz, threads=list(threads))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(m2), **outs[0])
    print(name, m2)


def fixture_liquid_sdf(name, source_npz, key, scale, threads, solid_key=None):
    """ParticleLevelSet::_computeSignedDistanceFromParticles on the positions of another fixture (not stored again);
    radius = scale * _liquidSDFParticleRadius (scale 2 = the smooth surface-tension kernel, fluidsimulation.cpp:5594-5597).
    The result is a minimum over particles: it must not depend on the reference's thread count."""
    z = np.load(os.path.join(OUT, source_npz + ".npz"))
    meta = json.loads(str(z["meta"]))
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    radius = float(scale * 0.5 * dx * np.sqrt(3.0))
    outs = []
    for t in threads:
        d = tempfile.mkdtemp(prefix="ffgold_")
        save_inputs(d, pos=z[key], phi=None if solid_key is None else z[solid_key])
        run("liquidsdf", d, I=I, J=J, K=K, dx=float(dx), radius=radius, threads=t)
        outs.append(np.load(os.path.join(d, "out_phi.npy")))
        post = np.load(os.path.join(d, "out_phi_post.npy")) if solid_key is not None else None    # + postProcessSignedDistanceField
        shutil.rmtree(d)
    for o in outs[1:]:
        assert o.tobytes() == outs[0].tobytes(), "reference liquid SDF depends on the thread count"
    m2 = dict(I=I, J=J, K=K, dx=dx, radius=radius, source=source_npz, key=key, threads=list(threads), solid_key=solid_key)
    extra = {} if post is None else dict(out_phi_post=post)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(m2), out_phi=outs[0], **extra)
    print(name, m2, "cells below the far value:", int((outs[0] < np.float32(3.0 * dx)).sum()))


def fixture_attribute(name, source_npz, key, radius_cells, seed, threads):
    """AttributeToGridTransfer<float>::transfer of a random per-particle scalar on the positions of another fixture
    (not stored again; the attribute is regenerated from the seed), cell-centred grid, offset dx/2."""
    z = np.load(os.path.join(OUT, source_npz + ".npz"))
    meta = json.loads(str(z["meta"]))
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    pos = z[key]
    attr = (np.random.default_rng(seed).random(len(pos)) * 10.0).astype(np.float32)
    attr3 = np.random.default_rng(seed + 1000).random((len(pos), 3)).astype(np.float32)    # the vec3 flavour (colour), same positions
    radius = float(np.float32(radius_cells) * dx)            # float radius = _ageAttributeRadius * _dx; (fluidsimulation.cpp:6997)
    outs = []
    for t in threads:
        d = tempfile.mkdtemp(prefix="ffgold_")
        save_inputs(d, pos=pos, attr=attr, attr3=attr3)
        run("attribute", d, I=I, J=J, K=K, dx=float(dx), radius=radius, threads=t)
        outs.append((np.load(os.path.join(d, "out_grid.npy")), np.load(os.path.join(d, "out_valid.npy")).astype(np.uint8),
                     np.load(os.path.join(d, "out_grid3.npy")).reshape(K, J, I, 3), np.load(os.path.join(d, "out_valid3.npy")).astype(np.uint8)))
        shutil.rmtree(d)
    for o in outs[1:]:
        assert all(a.tobytes() == b.tobytes() for a, b in zip(o, outs[0])), "thread-count dependent"
    assert outs[0][1].tobytes() == outs[0][3].tobytes()       # one weight field, one mask
    m2 = dict(I=I, J=J, K=K, dx=dx, radius=radius, source=source_npz, key=key, seed=seed, threads=list(threads))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(m2), out_grid=outs[0][0], out_valid=outs[0][1],
                        out_grid3=outs[0][2])
    print(name, m2, "valid cells:", int(outs[0][1].sum()))


def fixture_remove(name, advect_npz, seed, open_mask=0, open_width=2):
    """_removeMarkerParticles on post-advection positions (reference-built solid SDF of the advect fixture):
    extra particles inside the obstacle, one cell crowded beyond the 250 cap, a few extreme velocities."""
    z = np.load(os.path.join(OUT, advect_npz + ".npz"))
    meta = json.loads(str(z["meta"]))
    I, J, K, dx, dt = meta["I"], meta["J"], meta["K"], meta["dx"], meta["dt"]
    rng = np.random.default_rng(seed)
    pos = z["out_pos"].copy()
    inside = (np.array([0.12, 0.05, 0.11]) + rng.normal(0, 0.012, (400, 3))).astype(np.float32)     # around the sphere obstacle
    crowd = ((np.array([9.0, 8.0, 10.0]) + rng.random((330, 3))) * dx).astype(np.float32)          # one cell, > 250
    pos = np.concatenate([pos[: len(pos) // 2], inside, pos[len(pos) // 2:], crowd])
    pos = pos[rng.permutation(len(pos))]                                                           # index order matters
    vel = (rng.standard_normal(pos.shape) * 0.4).astype(np.float32)
    fast = rng.choice(len(pos), 9, replace=False)
    vel[fast] *= np.array([40, 45, 50, 55, 60, 300, 310, 320, 2000], np.float32)[:, None]          # histogram tail + outliers
    d = tempfile.mkdtemp(prefix="ffgold_")
    save_inputs(d, pos=pos, vel=vel, phi=z["in_phi"])
    info = run("remove", d, I=I, J=J, K=K, dx=float(dx), dt=float(dt), cfl=5, open=open_mask, open_width=open_width)
    out_pos, out_vel = np.load(os.path.join(d, "out_pos.npy")), np.load(os.path.join(d, "out_vel.npy"))
    shutil.rmtree(d)
    m2 = dict(I=I, J=J, K=K, dx=dx, dt=dt, cfl=5.0, particles=int(len(pos)), survivors=int(len(out_pos)), extreme=int(info["extreme"]),
              open_mask=open_mask)
    # the six planes of fluidsimulation.cpp:7780-7788 from the reference's own boundary box: float(min) + float(width * dx)
    buf = np.float32(open_width * dx)
    lo, hi = np.array(info["box_min"], np.float32) + buf, np.array(info["box_max"], np.float32) - buf
    planes = np.array([lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]], np.float32)
    closed = np.array([-np.inf, np.inf] * 3, np.float32)
    bounds = np.where([(open_mask >> q) & 1 for q in range(6)], planes, closed).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(m2), in_pos=pos, in_vel=vel, in_phi=z["in_phi"],
                        in_bounds=bounds, out_pos=out_pos, out_vel=out_vel)
    print(name, m2)


if __name__ == "__main__":
    if not os.path.exists(HARNESS):
        sys.exit("build oracle/_ref first: make -C oracle -j8 all")
    # whole-substep chains (reference-produced velocities, SDF, near-solid grid)
    fixture_scene("scene_flip_24x20x22_nondyadic", 24, 20, 22, 0.01, "flip", 2, "sphere:0.12,0.05,0.11,0.035", 1 / 60, 11)
    fixture_scene("scene_apic_22x24x20_dyadic", 22, 24, 20, 1.0 / 16.0, "apic", 1, "none", 1 / 60, 12)
    # stage-level P2G with seam-adversarial particles
    fixture_p2g("p2g_flip_23x21x25_seams", 23, 21, 25, 0.004, "flip", 21)
    fixture_p2g("p2g_apic_23x21x25_seams", 23, 21, 25, 0.004, "apic", 22)
    fixture_p2g("p2g_apic_20x20x20_dyadic", 20, 20, 20, 0.125, "apic", 23)
    fixture_p2g("p2g_flip_21x20x22_radius2", 21, 20, 22, 0.01, "flip", 24, radius_scale=2.0)
    # collision-heavy advection
    fixture_advect("advect_collide_24x20x22", "scene_flip_24x20x22_nondyadic", 31)
    # valid-face extrapolation of the reference's own P2G output (inputs: s1_* of the scene fixtures)
    fixture_extrapolate("extrapolate_flip_24x20x22", "scene_flip_24x20x22_nondyadic", 12, (1, 3, 16))
    fixture_extrapolate("extrapolate_apic_22x24x20", "scene_apic_22x24x20_dyadic", 12, (1, 3, 16))
    # marker-particle removal (oracle groundwork for the next row f2)
    fixture_remove("remove_24x20x22", "advect_collide_24x20x22", 41)
    fixture_remove("remove_open_24x20x22", "advect_collide_24x20x22", 43, open_mask=2 | 16 | 8)      # x+, z-, y+ open
    # liquid SDF from particles (oracle groundwork for row f3)
    fixture_liquid_sdf("liquid_sdf_23x21x25_seams", "p2g_flip_23x21x25_seams", "in_pos", 1.0, (1, 3, 16))
    fixture_liquid_sdf("liquid_sdf_22x24x20_radius2", "scene_apic_22x24x20_dyadic", "s0_pos", 2.0, (1, 16))
    fixture_liquid_sdf("liquid_sdf_post_24x20x22", "remove_24x20x22", "in_pos", 1.0, (1, 16), solid_key="in_phi")
    # scalar attribute P2G (oracle groundwork for row f4)
    fixture_attribute("attribute_23x21x25_seams_r2", "p2g_flip_23x21x25_seams", "in_pos", 2.0, 9, (1, 16))
    fixture_attribute("attribute_24x20x22_r1", "scene_flip_24x20x22_nondyadic", "s0_pos", 1.0, 10, (1, 16))   # _ageAttributeRadius
