"""Full-size GPU checks (-m gpu) at BASELINE.json's configurations, where the single-threaded
oracle would take minutes: size-independent properties of the three stages.

  config #2  dam break 128^3, APIC, ~4.6 M particles
  config #3  fill box 256^3 with a sphere obstacle, FLIP, ~40 M particles, collisions active

Properties (each pins a different part of the arithmetic):
  * partition of unity: particles carrying one constant velocity (zero affine) must give exactly
    that constant on every valid face (normalised weights), and a face is valid iff a particle
    is within reach: valid faces form the dilated particle cells;
  * linearity of P2G in the particle velocities;
  * a linear MAC field is reproduced by the trilinear gathers: G2P returns the field at the
    particle, APIC rows return its gradient;
  * a constant field advects every non-colliding particle by exactly v*dt (RK3), colliding ones
    stay outside the solid and inside the boundary;
  * binning: keys ascending, cells consistent with positions, permutation is a bijection;
  * run-to-run determinism (bitwise).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from blender_flip_fluids_b200 import engine
    engine.load_library()
    return engine


def _linear_field(I, J, K, dx, g, c):
    """MAC samples of v(x) = c + G x at the face centres (float64 -> float32)."""
    out = []
    for d, shape in enumerate([(K, J, I + 1), (K, J + 1, I), (K + 1, J, I)]):
        z, y, x = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), np.arange(shape[2]), indexing="ij", sparse=True)
        px = (x + (0.0 if d == 0 else 0.5)) * dx
        py = (y + (0.0 if d == 1 else 0.5)) * dx
        pz = (z + (0.0 if d == 2 else 0.5)) * dx
        out.append((c[d] + g[d][0] * px + g[d][1] * py + g[d][2] * pz).astype(np.float32))
    return out


def _check_binning(ctx, sc):
    cell, hkey, perm = ctx.get_binning()
    n = sc.n
    ci = np.floor(sc.pos.astype(np.float64) * (1.0 / sc.dx)).astype(np.int64)
    assert np.array_equal(cell, ci[:, 0] + sc.isize * (ci[:, 1] + sc.jsize * ci[:, 2]))
    ks = hkey[perm].astype(np.int64)
    assert (np.diff(ks) >= 0).all()
    same = np.diff(ks) == 0
    assert (np.diff(perm.astype(np.int64))[same] > 0).all()           # ties: ascending original index
    assert np.array_equal(np.bincount(perm, minlength=n), np.ones(n, np.int64))


def _run_properties(eng, sc, method, phi, near, ratio):
    I, J, K, dx = sc.isize, sc.jsize, sc.ksize, sc.dx
    apic = method == eng.APIC
    n = sc.n
    zero = np.zeros((n, 3), np.float32)
    const = np.array([0.37, -1.21, 0.58], np.float32)
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_solid(phi, near)
        # ---- binning
        ctx.set_particles(sc.pos, sc.vel, *( [sc.affx, sc.affy, sc.affz] if apic else [None] * 3))
        _check_binning(ctx, sc)
        # ---- P2G: determinism + linearity + partition of unity
        ctx.p2g(sc.radius, method)
        (u1, v1, w1), (m1u, m1v, m1w) = ctx.get_velocity_field()
        ctx.set_particles(sc.pos, sc.vel, *( [sc.affx, sc.affy, sc.affz] if apic else [None] * 3))
        ctx.p2g(sc.radius, method)
        (u1b, v1b, w1b), (m1ub, _, _) = ctx.get_velocity_field()
        assert u1.tobytes() == u1b.tobytes() and v1.tobytes() == v1b.tobytes() and w1.tobytes() == w1b.tobytes()
        assert m1u.tobytes() == m1ub.tobytes()
        vel2 = np.tile(const, (n, 1))
        ctx.set_particles(sc.pos, vel2, *( [zero, zero, zero] if apic else [None] * 3))
        ctx.p2g(sc.radius, method)
        (u2, v2, w2), (m2u, m2v, m2w) = ctx.get_velocity_field()
        assert np.array_equal(m1u, m2u) and np.array_equal(m1v, m2v) and np.array_equal(m1w, m2w)   # masks ignore velocities
        for f, m, c in ((u2, m2u, const[0]), (v2, m2v, const[1]), (w2, m2w, const[2])):
            assert m.sum() > 0
            assert np.abs(f[m == 1] - c).max() <= 1e-5 * abs(c)
            assert np.abs(f[m == 0]).max() <= 1e-5 * abs(c)           # un-normalised leftovers only (sum w <= 1e-6)
        # linearity: P2G(vel + 2*const) = P2G(vel) + 2*const on valid faces (same affine)
        ctx.set_particles(sc.pos, sc.vel + 2 * const, *( [sc.affx, sc.affy, sc.affz] if apic else [None] * 3))
        ctx.p2g(sc.radius, method)
        (u3, v3, w3), _ = ctx.get_velocity_field()
        scale = float(max(np.abs(u1).max(), np.abs(v1).max(), np.abs(w1).max())) + 2 * float(np.abs(const).max())
        for a, b, m, c in ((u3, u1, m1u, const[0]), (v3, v1, m1v, const[1]), (w3, w1, m1w, const[2])):
            assert np.abs(a[m == 1] - (b[m == 1] + 2 * c)).max() <= 2e-5 * scale
        # valid faces = faces with a particle in reach: every cell holding a particle has all its faces valid
        ci = np.floor(sc.pos.astype(np.float64) * (1.0 / dx)).astype(np.int64)
        occ = np.zeros((K, J, I), bool)
        occ[ci[:, 2], ci[:, 1], ci[:, 0]] = True
        if apic:      # the trilinear tent of any particle in a cell reaches all six faces of that cell
            assert m1u[:, :, :-1][occ].all() and m1u[:, :, 1:][occ].all()
            assert m1v[:, :-1, :][occ].all() and m1w[:-1, :, :][occ].all()
        # ---- G2P reproduces a linear field
        g = [[0.8, -0.3, 0.5], [0.2, 0.9, -0.6], [-0.7, 0.4, 0.1]]
        c = [0.3, -0.2, 0.1]
        mac = _linear_field(I, J, K, dx, g, c)
        ctx.set_particles(sc.pos, sc.vel, *( [sc.affx, sc.affy, sc.affz] if apic else [None] * 3))
        ctx.set_velocity_field(*mac)
        ctx.set_velocity_field(*mac, saved=True)
        ctx.sort_particles()
        ctx.g2p(method, ratio)
        _, vel_out, ax, ay, az = ctx.get_particles(pos=False, vel=True, affine=apic)
        p64 = sc.pos.astype(np.float64)
        want = np.stack([c[d] + p64 @ np.array(g[d]) for d in range(3)], axis=1)
        vscale = np.abs(want).max()
        if apic:
            assert np.abs(vel_out - want).max() <= 1e-5 * vscale
            for d, a in enumerate((ax, ay, az)):
                assert np.abs(a - np.array(g[d], np.float32)).max() <= 2e-4      # float gradient weights / dx
        else:
            # saved == current field: vFLIP = v_old, v = r*vPIC + (1-r)*v_old
            want_flip = ratio * want + (1 - ratio) * sc.vel.astype(np.float64)
            assert np.abs(vel_out - want_flip).max() <= 1e-5 * max(vscale, np.abs(sc.vel).max())
        # ---- advection through a constant field
        vconst = [0.21, -0.33, 0.17]
        macc = [np.full(m.shape, vconst[d], np.float32) for d, m in enumerate(mac)]
        ctx.set_particles(sc.pos, sc.vel)
        ctx.set_velocity_field(*macc)
        dt = 1.7 * dx / 0.33
        ctx.advect(dt, 5.0, False)
        free, *_ = ctx.get_particles(pos=True, vel=False)
        step = np.array(vconst) * dt
        interior = ((p64 > 2 * dx) & (p64 + step > 2 * dx) & (p64 < (np.array([I, J, K]) - 2) * dx)).all(axis=1)
        assert np.abs(free[interior] - (p64[interior] + step)).max() <= 4e-7 * max(I, J, K) * dx + 1e-5 * np.abs(step).max()
        ctx.set_particles(sc.pos, sc.vel)
        ctx.advect(dt, 5.0, True)
        coll, *_ = ctx.get_particles(pos=True, vel=False)
        moved = (coll != free).any(axis=1)
        lo = 1.5 * dx + 5e-5 + 0.1 * dx - 1e-6
        hi = np.array([I, J, K]) * dx - lo
        assert (coll >= lo - 1e-6).all() and (coll <= hi + 1e-6).all()   # inside the collision boundary
        return int(moved.sum())


def test_config2_dam_break_128_apic(eng):
    from blender_flip_fluids_b200 import scenes
    sc = scenes.dam_break(128, apic=True, vel="random", v0=0.5, seed=1234)
    phi, near = scenes.analytic_solid_sdf(128, 128, 128, sc.dx)
    _run_properties(eng, sc, eng.APIC, phi, near, 0.05)


def test_config3_fill_box_256_flip_obstacle(eng):
    from blender_flip_fluids_b200 import scenes
    sc = scenes.fill_box(256, apic=False, vel="random", v0=0.5, seed=77)
    n = 256
    sphere = (0.5, 0.25, 0.5, 0.15)
    phi, near = scenes.analytic_solid_sdf(n, n, n, sc.dx, sphere=sphere)
    # particles inside the obstacle are removed by the reference before the path runs
    d = np.linalg.norm(sc.pos.astype(np.float64) - np.array(sphere[:3]), axis=1)
    keep = d > sphere[3] + 0.5 * sc.dx
    sc.pos, sc.vel = sc.pos[keep], sc.vel[keep]
    assert sc.n > 30_000_000
    moved = _run_properties(eng, sc, eng.FLIP, phi, near, 0.02)
    assert moved > 0                                     # the collision projection really fired


@pytest.mark.parametrize("method_name", ["flip", "apic"])
def test_pipelined_host_field_io_matches_device_path(eng, method_name):
    """The host-buffer entry points move a large field in pieces under the kernels (ffb200_velocity_advector_advect:
    each direction's faces and masks leave behind that direction's kernels; ffb200_update_marker_particle_velocities:
    the field arrives in plane chunks, the gather follows range by range over the sorted particles). Same bits as the
    plain sequence set field -> kernel -> get, on a random (not smooth) field so that a particle launched before its
    planes arrived cannot go unnoticed; some particles sit outside the grid in z (clamped bins, zero velocity)."""
    from blender_flip_fluids_b200 import scenes
    method = eng.APIC if method_name == "apic" else eng.FLIP
    apic = method == eng.APIC
    sc = scenes.dam_break(128, apic=apic, vel="random", v0=0.5, seed=4321)
    rng = np.random.default_rng(99)
    pos = sc.pos.copy()
    pos[:50, 2] = -0.5 * sc.dx                                  # below the grid
    pos[50:100, 2] = (128 + 0.5) * sc.dx                         # above it
    aff = [sc.affx, sc.affy, sc.affz] if apic else [None] * 3
    mac = [rng.standard_normal(s).astype(np.float32) for s in eng.mac_shapes(128, 128, 128)]
    saved = [rng.standard_normal(s).astype(np.float32) for s in eng.mac_shapes(128, 128, 128)]
    with eng.FlipContext(128, 128, 128, sc.dx) as ctx:
        ctx.set_particles(pos, sc.vel, *aff)
        ctx.p2g(sc.radius, method)
        f_want, m_want = ctx.get_velocity_field()
        ctx.set_velocity_field(*mac)
        ctx.set_velocity_field(*saved, saved=True)
        ctx.sort_particles()
        ctx.g2p(method, 0.05)
        _, v_want, *a_want = ctx.get_particles(pos=False, vel=True, affine=apic)
    with eng.FlipContext(128, 128, 128, sc.dx) as ctx:
        f_got, m_got = ctx.velocity_advector_advect(pos, sc.vel, *aff, radius=sc.radius, method=method)
        for a, b in zip(list(f_want) + list(m_want), list(f_got) + list(m_got)):
            assert a.tobytes() == b.tobytes()
        got = ctx.update_marker_particle_velocities(pos, sc.vel, mac, saved=None if apic else saved, method=method, ratio_pic_flip=0.05)
        if apic:
            assert got[0].tobytes() == v_want.tobytes()
            for a, b in zip(a_want, got[1:]):
                assert a.tobytes() == b.tobytes()
        else:
            assert got.tobytes() == v_want.tobytes()
        # resident particles (the drop-in's protocol): the same call with null particle pointers
        ctx.declare_resident(particles=True)
        ctx.update_marker_particle_velocities(None, None, mac, saved=None if apic else saved, method=method, ratio_pic_flip=0.05)
        _, v_res, *a_res = ctx.get_particles(pos=False, vel=True, affine=apic)
        if apic:       # APIC output does not depend on the old velocity; FLIP's does (v_old was just replaced)
            assert v_res.tobytes() == v_want.tobytes()
            for a, b in zip(a_want, a_res):
                assert a.tobytes() == b.tobytes()
