"""Full-size parity against the UNMODIFIED reference (-m gpu): BASELINE.json configs[1] in full and a
21 M-particle slab of configs[2], stage by stage through oracle/_ref/ref_harness (the reference's own
VelocityAdvector::advect, _updateMarkerParticleVelocitiesThread and _advanceMarkerParticlesThread, built from
/root/reference by oracle/Makefile and shipped with the tree) on the same arrays the CUDA path gets.

Bars (north star): cell binning, sort order and valid-face masks bit-exact; grid velocities within 1e-5
(relative to max(|ref|, max|ref|), the tolerance of tests/test_gpu_parity.py); G2P velocities, APIC rows and
advected positions are asserted BIT-exact (stronger than the 1e-5 asked), since those kernels repeat the
reference arithmetic. The G2P and the advection take the reference's own P2G output as their field, so the
comparison is stage-level on identical inputs.
"""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


@pytest.fixture(scope="module")
def eng():
    from blender_flip_fluids_b200 import engine
    engine.load_library()
    return engine


def _harness(mode, d, **kv):
    assert os.path.exists(HARNESS), "oracle/_ref/ref_harness is missing (run __graft_entry__.build() where /root/reference exists)"
    r = subprocess.run([HARNESS, mode, d] + [f"{k}={v!r}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items()],
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _save(d, **arrs):
    for k, a in arrs.items():
        np.save(os.path.join(d, f"in_{k}.npy"), a)


def _load(d, name):
    return np.load(os.path.join(d, f"out_{name}.npy"))


def _close(a, b, tol=1e-5):
    b64 = b.astype(np.float64)
    scale = float(np.abs(b64).max())
    return bool(np.all(np.abs(a.astype(np.float64) - b64) <= tol * np.maximum(np.abs(b64), scale)))


def _stage_parity(eng, I, J, K, dx, pos, vel, aff, phi, near, method, ratio, dt):
    apic = method == "apic"
    m = eng.APIC if apic else eng.FLIP
    radius = 0.5 * dx * np.sqrt(3.0)
    common = dict(I=I, J=J, K=K, dx=float(dx), method=method)
    n = pos.shape[0]
    with tempfile.TemporaryDirectory(prefix="ffb200_full_") as d, eng.FlipContext(I, J, K, dx) as ctx:
        _save(d, pos=pos, vel=vel)
        if apic:
            _save(d, affx=aff[0], affy=aff[1], affz=aff[2])
        # ---- P2G ---------------------------------------------------------------------------------------------
        _harness("p2g", d, **common)
        ru, rv, rw = _load(d, "u"), _load(d, "v"), _load(d, "w")
        rmask = [_load(d, "validu"), _load(d, "validv"), _load(d, "validw")]
        ctx.set_solid(phi, near)
        ctx.set_particles(pos, vel, *(aff if apic else [None] * 3))
        # binning and sort order: Grid3d::positionToGridIndex cells, keys ascending, ties by ascending original index
        cell, hkey, perm = ctx.get_binning()
        ci = np.floor(pos.astype(np.float64) * (1.0 / dx)).astype(np.int64)
        inside = ((ci >= 0) & (ci < np.array([I, J, K]))).all(axis=1)
        assert np.array_equal(cell[inside], (ci[:, 0] + I * (ci[:, 1] + J * ci[:, 2]))[inside]) and (cell[~inside] == -1).all()
        ks = hkey[perm].astype(np.int64)
        assert (np.diff(ks) >= 0).all() and (np.diff(perm.astype(np.int64))[np.diff(ks) == 0] > 0).all()
        ctx.p2g(radius, m)
        (u, v, w), masks = ctx.get_velocity_field()
        for got, want, name in zip(masks, rmask, "uvw"):
            assert np.array_equal(got.astype(bool).ravel(), want.astype(bool).ravel()), f"valid mask {name} differs from the reference"
        for got, want, name in zip((u, v, w), (ru, rv, rw), "uvw"):
            assert _close(got.ravel(), want.ravel()), f"grid velocity {name} beyond 1e-5 of the reference"
        # ---- G2P on the reference's own field (FLIP: a second, shifted field as the saved one) ---------------------
        field = [np.ascontiguousarray(a.reshape(s)) for a, s in zip((ru, rv, rw), eng.mac_shapes(I, J, K))]
        _save(d, u=field[0], v=field[1], w=field[2])
        saved = None
        if not apic:
            saved = [np.ascontiguousarray(np.roll(f, 1, axis=2) * np.float32(0.9)) for f in field]
            _save(d, su=saved[0], sv=saved[1], sw=saved[2])
        _harness("g2p", d, ratio=float(ratio), **common)
        ctx.set_particles(pos, vel, *(aff if apic else [None] * 3))
        ctx.set_velocity_field(*field)
        if saved:
            ctx.set_velocity_field(*saved, saved=True)
        ctx.sort_particles()
        ctx.g2p(m, ratio)
        _, gvel, ax, ay, az = ctx.get_particles(pos=False, vel=True, affine=apic)
        assert gvel.tobytes() == _load(d, "vel").astype(np.float32).tobytes(), "G2P velocities are not bit-identical"
        if apic:
            for got, name in ((ax, "affx"), (ay, "affy"), (az, "affz")):
                assert got.tobytes() == _load(d, name).astype(np.float32).tobytes(), f"G2P {name} is not bit-identical"
        # ---- advection (RK3 + collision) through the same field ----------------------------------------------------
        _save(d, phi=phi, near=near)
        _harness("advect", d, dt=float(dt), cfl=5, **common)
        ctx.advect(dt, 5.0, True)
        gpos, *_ = ctx.get_particles(pos=True, vel=False)
        rpos = _load(d, "pos").astype(np.float32)
        assert gpos.shape == rpos.shape == (n, 3)
        moved = int((rpos != pos).any(axis=1).sum())
        assert gpos.tobytes() == rpos.tobytes(), "advected positions are not bit-identical"
        return moved


def test_config2_dam_break_128_apic_vs_reference(eng):
    """BASELINE configs[1] in full: dam break 128^3, APIC, 4 637 952 particles, RK3 + collision."""
    from blender_flip_fluids_b200 import scenes
    sc = scenes.dam_break(128, apic=True, vel="random", v0=0.5, seed=1234)
    phi, near = scenes.analytic_solid_sdf(128, 128, 128, sc.dx)
    assert sc.n == 4637952
    _stage_parity(eng, 128, 128, 128, sc.dx, sc.pos, sc.vel, [sc.affx, sc.affy, sc.affz], phi, near, "apic", 0.05, 1.0 * sc.dx / 0.5)


def test_config3_fill_box_slab_flip_obstacle_vs_reference(eng):
    """A 21 M-particle slab of BASELINE configs[2]: fill box 256 x 256 x 140 (dx = 1/256) with a sphere obstacle,
    FLIP 0.98 (PIC ratio 0.02), collision projection active."""
    from blender_flip_fluids_b200 import scenes
    I, J, K, dx = 256, 256, 140, 1.0 / 256
    rng = np.random.default_rng(77)
    pos = scenes._seed_cells(3, I - 3, 3, int(0.6 * J), 3, K - 3, dx, 8, rng)
    sphere = (0.5, 0.25, 0.27, 0.12)
    d = np.linalg.norm(pos.astype(np.float64) - np.array(sphere[:3]), axis=1)
    pos = np.ascontiguousarray(pos[d > sphere[3] + 0.5 * dx])      # particles inside the obstacle are removed before the path runs
    vel = (rng.uniform(-1.0, 1.0, size=pos.shape) * 0.5).astype(np.float32)
    assert pos.shape[0] > 20_000_000
    phi, near = scenes.analytic_solid_sdf(I, J, K, dx, sphere=sphere)
    moved = _stage_parity(eng, I, J, K, dx, pos, vel, None, phi, near, "flip", 0.02, 2.0 * dx / 0.5)
    assert moved > 0
