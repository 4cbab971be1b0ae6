"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libffb200.so via
blender_flip_fluids_b200.engine), against the C oracle on the same seeded inputs and against
the reference-generated golden fixtures in tests/golden/.

Bars (BASELINE.json north_star):
  * cell binning, sort order, valid-face masks: bit-exact;
  * grid velocities, particle velocities, APIC matrices, advected positions: <= 1e-5 relative.
    "Relative" is |a-b| <= RTOL * max(|b|, scale) with scale = the largest magnitude in the
    reference array (velocities near zero from cancellation are compared against the field's
    scale). G2P and advection repeat the reference's arithmetic exactly, so they are in fact
    asserted bit-exact below; the P2G sums differ only by summation order.
"""
import numpy as np
import pytest

from conftest import bits_equal, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(a, b, rtol=RTOL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = float(np.max(np.abs(b))) if b.size else 0.0
    err = np.abs(a - b)
    tol = rtol * np.maximum(np.abs(b), scale)
    ok = bool(np.all(err <= tol))
    if not ok:
        worst = np.unravel_index(np.argmax(err - tol), err.shape)
        print(f"close(): {int((err > tol).sum())} of {err.size} out of tolerance; worst at {worst}: got {a[worst]!r} "
              f"want {b[worst]!r} err {err[worst]:.3e} tol {tol[worst]:.3e} scale {scale:.3e}")
    return ok


@pytest.fixture(scope="module")
def eng():
    from blender_flip_fluids_b200 import engine
    engine.load_library()
    return engine


def _method(meta):
    return 1 if meta["method"] == "apic" else 0


P2G_FIXTURES = ["p2g_flip_23x21x25_seams", "p2g_apic_23x21x25_seams", "p2g_apic_20x20x20_dyadic",
                "p2g_flip_21x20x22_radius2"]
SCENES = ["scene_flip_24x20x22_nondyadic", "scene_apic_22x24x20_dyadic"]


@pytest.mark.parametrize("name", P2G_FIXTURES)
def test_p2g_golden(eng, name):
    meta, g = load_golden(name)
    aff = [g.get("in_aff" + c) for c in "xyz"]
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        (u, v, w), (vu, vv, vw) = ctx.velocity_advector_advect(g["in_pos"], g["in_vel"], *aff, radius=meta["radius"],
                                                               method=_method(meta))
    assert np.array_equal(vu, g["out_validu"]) and np.array_equal(vv, g["out_validv"]) and np.array_equal(vw, g["out_validw"])
    assert close(u, g["out_u"]) and close(v, g["out_v"]) and close(w, g["out_w"])


@pytest.mark.parametrize("name", P2G_FIXTURES)
def test_p2g_exact_path_is_bit_exact(eng, name):
    """Guard band forced open: every face is re-summed in the reference's order."""
    meta, g = load_golden(name)
    aff = [g.get("in_aff" + c) for c in "xyz"]
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        ctx.set_valid_guard(float("inf"), 0.0)
        (u, v, w), (vu, vv, vw) = ctx.velocity_advector_advect(g["in_pos"], g["in_vel"], *aff, radius=meta["radius"],
                                                               method=_method(meta))
    assert bits_equal(u, g["out_u"]) and bits_equal(v, g["out_v"]) and bits_equal(w, g["out_w"])
    assert np.array_equal(vu, g["out_validu"]) and np.array_equal(vv, g["out_validv"]) and np.array_equal(vw, g["out_validw"])


@pytest.mark.parametrize("name", SCENES)
def test_scene_chain_golden(eng, name):
    meta, g = load_golden(name)
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    apic = meta["method"] == "apic"
    aff = [g.get("s0_aff" + c) for c in "xyz"]
    mac = (g["s2_u"], g["s2_v"], g["s2_w"])
    with eng.FlipContext(I, J, K, dx) as ctx:
        (u, v, w), (vu, vv, vw) = ctx.velocity_advector_advect(g["s0_pos"], g["s0_vel"], *aff, radius=meta["radius"],
                                                               method=_method(meta))
        assert np.array_equal(vu, g["s1_validu"]) and np.array_equal(vv, g["s1_validv"]) and np.array_equal(vw, g["s1_validw"])
        assert close(u, g["s1_u"]) and close(v, g["s1_v"]) and close(w, g["s1_w"])
        if apic:
            vel, ax, ay, az = ctx.update_marker_particle_velocities(g["s0_pos"], g["s0_vel"], mac, method=eng.APIC)
            assert bits_equal(ax, g["s3_affx"]) and bits_equal(ay, g["s3_affy"]) and bits_equal(az, g["s3_affz"])
        else:
            vel = ctx.update_marker_particle_velocities(g["s0_pos"], g["s0_vel"], mac,
                                                        saved=(g["s2_su"], g["s2_sv"], g["s2_sw"]), method=eng.FLIP,
                                                        ratio_pic_flip=meta["ratio"])
        assert bits_equal(vel, g["s3_vel"])
        out = ctx.advance_marker_particles(g["s0_pos"], mac, g["s2_phi"], g["s2_near"], dt=meta["dt"], cfl=meta["cfl"])
        assert bits_equal(out, g["s4_pos"])


def test_advect_collision_golden(eng):
    meta, g = load_golden("advect_collide_24x20x22")
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        out = ctx.advance_marker_particles(g["in_pos"], (g["in_u"], g["in_v"], g["in_w"]), g["in_phi"], g["in_near"],
                                           dt=meta["dt"], cfl=meta["cfl"])
    assert bits_equal(out, g["out_pos"])


def test_binning_and_sort_order(eng, oracle):
    meta, g = load_golden("p2g_flip_23x21x25_seams")
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    pos = g["in_pos"].copy()
    pos[:5] = [[-1e-3, 0.01, 0.01], [0.01, (J + 1) * dx, 0.01], [0.01, 0.01, K * dx + 1], [(I + 0.5) * dx, 0.0, 0.0], [0, 0, 0]]
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_particles(pos, g["in_vel"])
        cell, hkey, perm = ctx.get_binning()
        # round trip: particles come back in the original order, bit-identical
        p2, v2, *_ = ctx.get_particles()
    ocell, ohkey, operm = oracle.bin_sort(I, J, K, dx, pos)
    assert np.array_equal(cell, ocell)
    assert np.array_equal(hkey, ohkey)
    assert np.array_equal(perm, operm)
    assert bits_equal(p2, pos) and bits_equal(v2, g["in_vel"])


@pytest.mark.parametrize("method", ["flip", "apic"])
@pytest.mark.parametrize("n,dx", [(32, 1.0 / 32), (30, 0.004 * 250 / 30)])
def test_substep_vs_oracle(eng, oracle, method, n, dx):
    """Dam break, all three stages against the oracle on identical inputs (dyadic and not)."""
    from blender_flip_fluids_b200 import scenes
    apic = method == "apic"
    sc = scenes.dam_break(n, apic=apic, dx=dx, vel="random", v0=0.5, seed=5)
    I = J = K = n
    m = eng.APIC if apic else eng.FLIP
    aff = (sc.affx, sc.affy, sc.affz)
    (ou, ov, ow), (ovu, ovv, ovw) = oracle.p2g(I, J, K, dx, sc.radius, m, sc.pos, sc.vel, *aff)
    phi, near = scenes.analytic_solid_sdf(I, J, K, dx, sphere=(0.3 * n * dx, 0.3 * n * dx, 0.5 * n * dx, 0.12 * n * dx))
    dt = 2.5 * dx / 0.5
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_particles(sc.pos, sc.vel, *aff)
        ctx.p2g(sc.radius, m)
        (u, v, w), (vu, vv, vw) = ctx.get_velocity_field()
        assert np.array_equal(vu, ovu) and np.array_equal(vv, ovv) and np.array_equal(vw, ovw)
        assert close(u, ou) and close(v, ov) and close(w, ow)
        # feed the ORACLE's grid to both sides so G2P/advect are compared on identical inputs
        ctx.set_velocity_field(ou, ov, ow)
        ctx.set_velocity_field(ou * 0.9, ov * 0.9, ow * 0.9, saved=True)
        ctx.set_solid(phi, near)
        ctx.g2p(m, 0.05)
        _, vel, ax, ay, az = ctx.get_particles(pos=False, vel=True, affine=apic)
        ctx.advect(dt, 5.0, True)
        pos1, *_ = ctx.get_particles(pos=True, vel=False)
    if apic:
        ovel, oax, oay, oaz = oracle.g2p_apic(I, J, K, dx, sc.pos, (ou, ov, ow))
        assert bits_equal(ax, oax) and bits_equal(ay, oay) and bits_equal(az, oaz)
    else:
        ovel = oracle.g2p_flip(I, J, K, dx, sc.pos, sc.vel, (ou, ov, ow), (ou * 0.9, ov * 0.9, ow * 0.9), 0.05)
    assert bits_equal(vel, ovel)
    opos = oracle.advect(I, J, K, dx, sc.pos, (ou, ov, ow), phi, near, dt, 5.0, True)
    assert bits_equal(pos1, opos)
    free = oracle.advect(I, J, K, dx, sc.pos, (ou, ov, ow), phi, near, dt, 5.0, False)
    assert (free != opos).any()


def test_empty_and_tiny_inputs(eng):
    with eng.FlipContext(12, 11, 13, 0.1) as ctx:
        z = np.zeros((0, 3), np.float32)
        (u, v, w), (vu, vv, vw) = ctx.velocity_advector_advect(z, z)
        assert not u.any() and not v.any() and not w.any() and not vu.any() and not vv.any() and not vw.any()
        p = np.array([[0.55, 0.52, 0.57]], np.float32)
        vel = np.array([[1.0, 2.0, 3.0]], np.float32)
        (u, v, w), (vu, vv, vw) = ctx.velocity_advector_advect(p, vel)
        assert vu.sum() > 0 and vv.sum() > 0 and vw.sum() > 0
        assert np.allclose(u[vu == 1], 1.0) and np.allclose(v[vv == 1], 2.0) and np.allclose(w[vw == 1], 3.0)


def test_apic_seam_drop_probe(eng):
    """SURVEY appendix A probe: N=32, dx=0.1, particle at (9.05dx, 6dx, 6dx): validU(9,5,5) is
    set, validU(10,5,5) is NOT (its 10^3 block never sees the 'simple' particle)."""
    dx = 0.1
    with eng.FlipContext(32, 32, 32, dx) as ctx:
        z = np.zeros((1, 3), np.float32)
        vel = np.array([[1.0, 2.0, 3.0]], np.float32)
        p = np.array([[9.05 * dx, 6 * dx, 6 * dx]], np.float32)
        (_, _, _), (vu, _, _) = ctx.velocity_advector_advect(p, vel, z, z, z, method=eng.APIC)
        assert vu[5, 5, 9] == 1 and vu[5, 5, 10] == 0
        p = np.array([[8.95 * dx, 6 * dx, 6 * dx]], np.float32)
        (_, _, _), (vu, _, _) = ctx.velocity_advector_advect(p, vel, z, z, z, method=eng.APIC)
        assert vu[5, 5, 8] == 1 and vu[5, 5, 9] == 1


def test_determinism(eng):
    from blender_flip_fluids_b200 import scenes
    sc = scenes.dam_break(24, apic=True, seed=9)
    outs = []
    for _ in range(2):
        with eng.FlipContext(24, 24, 24, sc.dx) as ctx:
            (u, v, w), _ = ctx.velocity_advector_advect(sc.pos, sc.vel, sc.affx, sc.affy, sc.affz, method=eng.APIC)
            outs.append((u, v, w))
    assert all(bits_equal(a, b) for a, b in zip(*outs))


VARIANT_CHECK = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import load_golden
from blender_flip_fluids_b200 import engine
for name in ("p2g_flip_23x21x25_seams", "p2g_apic_23x21x25_seams"):
    meta, g = load_golden(name)
    aff = [g.get("in_aff" + c) for c in "xyz"]
    m = 1 if meta["method"] == "apic" else 0
    for guard in (None, float("inf")):
        with engine.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
            if guard is not None:
                ctx.set_valid_guard(guard, 0.0)
            (u, v, w), (vu, vv, vw) = ctx.velocity_advector_advect(g["in_pos"], g["in_vel"], *aff, radius=meta["radius"], method=m)
            cell, hkey, perm = ctx.get_binning()
        assert np.array_equal(vu, g["out_validu"]) and np.array_equal(vv, g["out_validv"]) and np.array_equal(vw, g["out_validw"])
        for a, b in ((u, g["out_u"]), (v, g["out_v"]), (w, g["out_w"])):
            if guard is None:
                s = np.abs(b).max()
                assert np.all(np.abs(a.astype(np.float64) - b) <= 1e-5 * np.maximum(np.abs(b), s))
            else:
                assert a.tobytes() == b.tobytes()
        ks = hkey[perm].astype(np.int64)
        assert (np.diff(ks) >= 0).all() and (np.diff(perm.astype(np.int64))[np.diff(ks) == 0] > 0).all()
# resident G2P -> advection on the reference's own fields (bit-exact with and without the stage-1 reuse)
for name in ("scene_flip_24x20x22_nondyadic", "scene_apic_22x24x20_dyadic"):
    meta, g = load_golden(name)
    apic = meta["method"] == "apic"
    aff = [g.get("s0_aff" + c) for c in "xyz"]
    with engine.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        ctx.set_solid(g["s2_phi"], g["s2_near"])
        ctx.set_particles(g["s0_pos"], g["s0_vel"], *aff)
        ctx.set_velocity_field(g["s2_u"], g["s2_v"], g["s2_w"])
        ctx.set_velocity_field(g["s2_su"], g["s2_sv"], g["s2_sw"], saved=True)
        ctx.sort_particles()
        ctx.g2p(1 if apic else 0, meta["ratio"])
        ctx.advect(meta["dt"], meta["cfl"], True)
        pos, vel, *_ = ctx.get_particles(pos=True, vel=True, affine=apic)
    assert vel.tobytes() == g["s3_vel"].tobytes() and pos.tobytes() == g["s4_pos"].tobytes()
print("ok")
'''


@pytest.mark.parametrize("env", [{"FFB200_P2G_VARIANT": "1"}, {"FFB200_P2G_VARIANT": "2"},
                                 {"FFB200_SORT": "radix"}, {"FFB200_P2G_STREAMS": "0", "FFB200_FUSE_SEAM": "0"},
                                 {"FFB200_REUSE_G2P": "0"}])
def test_alternate_kernel_paths(env, tmp_path):
    """The gather formulations of the P2G (brick: radii up to 2 dx; whole grid: the attribute transfer's kernel), the LSD radix sort and the switched-off fusions (single P2G
    stream, stand-alone membership pass, no RK3 stage-1 reuse) stay selectable and correct."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", VARIANT_CHECK, root], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


# ---- valid-face extrapolation (GridUtils::extrapolateGrid) ----------------------------------------------------
@pytest.mark.parametrize("name,scene", [("extrapolate_flip_24x20x22", "scene_flip_24x20x22_nondyadic"),
                                        ("extrapolate_apic_22x24x20", "scene_apic_22x24x20_dyadic")])
def test_extrapolate_golden(eng, name, scene):
    """Device extrapolation against the unmodified reference's output: bit-exact."""
    meta, e = load_golden(name)
    _, g = load_golden(scene)
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        ctx.set_velocity_field(g["s1_u"], g["s1_v"], g["s1_w"])
        ctx.set_valid_velocities(g["s1_validu"], g["s1_validv"], g["s1_validw"])
        ctx.extrapolate_velocity_field(meta["layers"])
        (u, v, w), _ = ctx.get_velocity_field()
        assert bits_equal(u, e["out_u"]) and bits_equal(v, e["out_v"]) and bits_equal(w, e["out_w"])
        # default layer count = the reference's ceil(sqrt(3) * CFL) + 3
        ctx.set_velocity_field(g["s1_u"], g["s1_v"], g["s1_w"])
        ctx.extrapolate_velocity_field()
        (u2, _, _), _ = ctx.get_velocity_field()
        assert bits_equal(u2, e["out_u"])


@pytest.mark.parametrize("layers", [0, 1, 5, 12])
def test_extrapolate_vs_oracle_random_masks(eng, oracle, layers):
    rng = np.random.default_rng(100 + layers)
    I, J, K, dx = 19, 17, 23, 0.013
    shapes = eng.mac_shapes(I, J, K)
    grids = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    masks = [(rng.random(s) < p).astype(np.uint8) for s, p in zip(shapes, (0.02, 0.3, 0.0005))]
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_velocity_field(*grids)
        ctx.set_valid_velocities(*masks)
        ctx.extrapolate_velocity_field(layers)
        got, _ = ctx.get_velocity_field()
    for a, gr, m in zip(got, grids, masks):
        assert bits_equal(a, oracle.extrapolate(gr, m, layers))


def test_p2g_extrapolate_save_chain(eng, oracle):
    """The reference's 'Advect Velocity Field' stage on the device: P2G -> extrapolate -> save
    (fluidsimulation.cpp:5652-5654, 5671-5679), bit-exact in the exact-sum P2G mode."""
    meta, g = load_golden("scene_apic_22x24x20_dyadic")
    _, e = load_golden("extrapolate_apic_22x24x20")
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        ctx.set_valid_guard(float("inf"), 0.0)
        ctx.set_particles(g["s0_pos"], g["s0_vel"], g["s0_affx"], g["s0_affy"], g["s0_affz"])
        ctx.p2g(meta["radius"], eng.APIC)
        ctx.extrapolate_velocity_field()
        ctx.save_velocity_field()
        (u, v, w), _ = ctx.get_velocity_field()
        assert bits_equal(u, e["out_u"]) and bits_equal(v, e["out_v"]) and bits_equal(w, e["out_w"])


@pytest.mark.parametrize("method", ["flip", "apic"])
def test_declare_resident_matches_full_uploads(eng, method):
    """ffb200_declare_resident only skips uploads: the three host-buffer entry points give the same
    bytes with and without it (and a wrong particle count is refused)."""
    from blender_flip_fluids_b200 import scenes
    apic = method == "apic"
    m = eng.APIC if apic else eng.FLIP
    sc = scenes.dam_break(24, apic=apic, vel="random", v0=0.4, seed=9)
    phi, near = scenes.analytic_solid_sdf(24, 24, 24, sc.dx)
    aff = [sc.affx, sc.affy, sc.affz] if apic else [None] * 3
    results = []
    for declare in (False, True):
        with eng.FlipContext(24, 24, 24, sc.dx) as ctx:
            pos, vel = sc.pos.copy(), sc.vel.copy()
            (u, v, w), _ = ctx.velocity_advector_advect(pos, vel, *aff, radius=sc.radius, method=m)
            saved = None if apic else (u * 0.5, v * 0.5, w * 0.5)
            if declare:
                ctx.declare_resident(particles=True)
            g2p = ctx.update_marker_particle_velocities(pos, vel, (u, v, w), saved=saved, method=m, ratio_pic_flip=0.05)
            if declare:
                ctx.declare_resident(particles=True, field=True)
            newpos = ctx.advance_marker_particles(pos, (u, v, w), phi, near, dt=0.8 * sc.dx / 0.4, cfl=5.0)
            results.append((g2p, newpos))
            if declare:
                ctx.declare_resident(particles=True)
                with pytest.raises(RuntimeError):
                    ctx.advance_marker_particles(pos[:-1], (u, v, w), phi, near, dt=0.01, cfl=5.0)
    (g0, p0), (g1, p1) = results
    if apic:
        assert all(bits_equal(a, b) for a, b in zip(g0, g1))
    else:
        assert bits_equal(g0, g1)
    assert bits_equal(p0, p1)
    assert (p0 != sc.pos).any()


@pytest.mark.parametrize("method", ["flip", "apic"])
def test_cuda_graph_capture_matches_eager(eng, method):
    """Whole substeps (sort, P2G, extrapolation, save, G2P, advection) captured into a CUDA graph and
    replayed give the same bytes as eager launches: the resident stages make no host decisions and
    skip their timing events while the stream is capturing."""
    import torch
    from blender_flip_fluids_b200 import scenes
    apic = method == "apic"
    m = eng.APIC if apic else eng.FLIP
    sc = scenes.dam_break(24, apic=apic, vel="random", v0=0.4, seed=3)
    phi, near = scenes.analytic_solid_sdf(24, 24, 24, sc.dx)
    aff = [sc.affx, sc.affy, sc.affz] if apic else [None] * 3
    dt = 0.7 * sc.dx / 0.4

    def substep(ctx):
        ctx.p2g(sc.radius, m)
        ctx.extrapolate_velocity_field()
        ctx.save_velocity_field()
        ctx.g2p(m, 0.05)
        ctx.advect(dt, 5.0, True)

    def run(graphed):
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            with eng.FlipContext(24, 24, 24, sc.dx) as ctx:
                ctx.set_stream(stream.cuda_stream)
                ctx.set_solid(phi, near)
                ctx.set_particles(sc.pos, sc.vel, *aff)
                substep(ctx)
                substep(ctx)                      # scratch buffers exist now; the double buffer is back at its start
                if graphed:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream):
                        substep(ctx)
                        substep(ctx)
                    g.replay()
                else:
                    substep(ctx)
                    substep(ctx)
                torch.cuda.synchronize()
                out = ctx.get_particles(pos=True, vel=True, affine=apic)
                field, _ = ctx.get_velocity_field()
                ctx.reset_stream()
        return out, field

    (p0, v0, *a0), f0 = run(False)
    (p1, v1, *a1), f1 = run(True)
    assert bits_equal(p0, p1) and bits_equal(v0, v1)
    assert all(bits_equal(x, y) for x, y in zip(f0, f1))
    if apic:
        assert all(bits_equal(x, y) for x, y in zip(a0, a1))
    assert (p0 != sc.pos).any()


@pytest.mark.parametrize("method", ["flip", "apic"])
def test_p2g_dense_and_clustered_cells_vs_oracle(eng, oracle, method):
    """27 particles per cell (several staging chunks per shifted cell) plus 600 particles crowded into
    one cell octant and 300 more on a block seam: the cell walk, the tie re-ranking of the sort and
    the seam frames with long runs."""
    from blender_flip_fluids_b200 import scenes
    apic = method == "apic"
    m = eng.APIC if apic else eng.FLIP
    n, dx = 22, 0.013
    sc = scenes.dam_break(n, ppc=27, apic=apic, dx=dx, vel="random", v0=0.7, seed=21)
    rng = np.random.default_rng(8)
    crowd = (np.array([7.1, 9.6, 12.3]) + rng.random((600, 3)) * 0.4) * dx          # one octant of cell (7, 9, 12)
    seam = (np.array([9.8, 10.2, 9.9]) + rng.random((300, 3)) * np.array([0.4, 0.1, 0.3])) * dx   # across the 10-node seams
    pos = np.concatenate([sc.pos, crowd.astype(np.float32), seam.astype(np.float32)])
    k = pos.shape[0] - sc.n
    vel = np.concatenate([sc.vel, rng.uniform(-1, 1, (k, 3)).astype(np.float32)])
    aff = [None] * 3
    if apic:
        aff = [np.concatenate([a, (rng.uniform(-1, 1, (k, 3)) * 0.1 / dx).astype(np.float32)]) for a in (sc.affx, sc.affy, sc.affz)]
    (ou, ov, ow), (ovu, ovv, ovw) = oracle.p2g(n, n, n, dx, sc.radius, m, pos, vel, *aff)
    with eng.FlipContext(n, n, n, dx) as ctx:
        ctx.set_particles(pos, vel, *aff)
        ctx.p2g(sc.radius, m)
        (u, v, w), (vu, vv, vw) = ctx.get_velocity_field()
        assert np.array_equal(vu, ovu) and np.array_equal(vv, ovv) and np.array_equal(vw, ovw)
        assert close(u, ou) and close(v, ov) and close(w, ow)
        cell, hkey, perm = ctx.get_binning()
        ocell, ohkey, operm = oracle.bin_sort(n, n, n, dx, pos)
        assert np.array_equal(cell, ocell) and np.array_equal(hkey, ohkey) and np.array_equal(perm, operm)
        ctx.set_valid_guard(float("inf"), 0.0)                 # every face through the reference-order sum
        ctx.set_particles(pos, vel, *aff)
        ctx.p2g(sc.radius, m)
        (u, v, w), _ = ctx.get_velocity_field()
        assert bits_equal(u, ou) and bits_equal(v, ov) and bits_equal(w, ow)


def test_maximum_particle_speed(eng, oracle):
    """_getMaximumMarkerParticleSpeed on the device == the oracle's (and numpy's float32 arithmetic), bit for bit."""
    rng = np.random.default_rng(17)
    pos = (rng.random((50001, 3)) * 0.8 + 0.1).astype(np.float32)
    vel = (rng.standard_normal((50001, 3)) * 3.0).astype(np.float32)
    with eng.FlipContext(16, 16, 16, 1.0 / 16) as ctx:
        ctx.set_particles(pos, vel)
        got = ctx.maximum_particle_speed()
        ctx.sort_particles()
        assert ctx.maximum_particle_speed() == got
        ctx.set_particles(pos[:0], vel[:0])
        assert ctx.maximum_particle_speed() == 0.0
    want = oracle.max_particle_speed(vel)
    d = (vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1]) + vel[:, 2] * vel[:, 2]
    assert got == want == float(np.sqrt(np.float64(d.max())))


@pytest.mark.parametrize("name", ["remove_24x20x22", "remove_open_24x20x22"])
def test_remove_particles_golden(eng, name):
    """_removeMarkerParticles on the device == the unmodified reference (survivors in order, bit for bit),
    whatever order the particles currently have on the device; closed domain and three open sides."""
    meta, g = load_golden(name)
    bounds = g["in_bounds"] if meta["open_mask"] else None
    for presort in (False, True):
        with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
            ctx.set_solid(g["in_phi"], np.zeros(ctx.near_dims, np.uint8))
            ctx.set_particles(g["in_pos"], g["in_vel"])
            if presort:
                ctx.sort_particles()
            remaining, extreme = ctx.remove_marker_particles(meta["dt"], meta["cfl"], open_bounds=bounds)
            p, v, *_ = ctx.get_particles()
        assert (remaining, extreme) == (meta["survivors"], meta["extreme"])
        assert bits_equal(p, g["out_pos"]) and bits_equal(v, g["out_vel"])


def test_mark_removed_host_entry_point(eng, oracle):
    """ffb200_mark_removed_marker_particles: the mask in the caller's order, with uploads and with everything declared
    resident after an advection-like sequence; open sides and a pre-removed (lifetime) mask against the oracle."""
    meta, g = load_golden("remove_open_24x20x22")
    I, J, K, dx, dt = meta["I"], meta["J"], meta["K"], meta["dx"], meta["dt"]
    pos, vel, phi = g["in_pos"], g["in_vel"], g["in_phi"]
    pre = (np.arange(len(pos)) % 5 == 0).astype(np.uint8)
    want = {}
    for key, kw in {"open": dict(open_bounds=g["in_bounds"]), "pre": dict(pre_removed=pre, max_per_cell=4)}.items():
        want[key] = oracle.remove_particles(I, J, K, dx, pos, vel, phi, dt, 5.0, **kw)
    with eng.FlipContext(I, J, K, dx) as ctx:
        near = np.zeros(ctx.near_dims, np.uint8)
        removed, extreme = ctx.mark_removed_marker_particles(pos, vel, phi, near, dt, open_bounds=g["in_bounds"])
        assert np.array_equal(removed, want["open"][0]) and extreme == want["open"][1] == meta["extreme"]
        assert bits_equal(pos[removed == 0], g["out_pos"])
        # resident: particles sorted on the device (what an advection leaves), nothing uploaded but the pre-removed mask
        ctx.sort_particles()
        ctx.declare_resident(particles=True, solid=True)
        removed, extreme = ctx.mark_removed_marker_particles(None, None, None, None, dt, max_particles_per_cell=4, pre_removed=pre)
        assert np.array_equal(removed, want["pre"][0]) and extreme == want["pre"][1]
        # a wrong resident claim fails loudly
        ctx.declare_resident(particles=True, solid=True)
        ctx.n += 1
        with pytest.raises(RuntimeError):
            ctx.mark_removed_marker_particles(None, None, None, None, dt)


@pytest.mark.parametrize("cap,extreme_on", [(250, True), (3, True), (1, False), (0, True)])
def test_remove_particles_vs_oracle(eng, oracle, cap, extreme_on):
    """Small per-cell caps put most cells over the cap (the order-exact ranking path); APIC rows travel with
    the survivors; the compacted particles feed the next P2G like freshly uploaded ones."""
    meta, g = load_golden("remove_24x20x22")
    I, J, K, dx, dt = meta["I"], meta["J"], meta["K"], meta["dx"], meta["dt"]
    rng = np.random.default_rng(100 + cap)
    n = 30011
    pos = (rng.random((n, 3)) * [I * dx, J * dx, K * dx]).astype(np.float32)
    pos[: n // 4] = g["in_pos"][: n // 4]
    pos[7] = [-0.5 * dx, 0.05, 0.05]                                    # outside the grid
    vel = (rng.standard_normal((n, 3)) * 0.5).astype(np.float32)
    vel[rng.choice(n, 11, replace=False)] *= np.float32(90.0)
    aff = [rng.standard_normal((n, 3)).astype(np.float32) for _ in range(3)]
    removed, want_extreme = oracle.remove_particles(I, J, K, dx, pos, vel, g["in_phi"], dt, 5.0, max_per_cell=cap,
                                                    extreme_removal=extreme_on)
    keep = removed == 0
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_solid(g["in_phi"], np.zeros(ctx.near_dims, np.uint8))
        ctx.set_particles(pos, vel, *aff)
        ctx.sort_particles()
        remaining, extreme = ctx.remove_marker_particles(dt, 5.0, max_particles_per_cell=cap, extreme_velocity_removal=extreme_on)
        p, v, ax, ay, az = ctx.get_particles(affine=True)
        assert (remaining, extreme) == (int(keep.sum()), want_extreme)
        assert bits_equal(p, pos[keep]) and bits_equal(v, vel[keep])
        assert bits_equal(ax, aff[0][keep]) and bits_equal(ay, aff[1][keep]) and bits_equal(az, aff[2][keep])
        if remaining:
            radius = 0.5 * dx * np.sqrt(3.0)
            ctx.p2g(radius, 1)
            (u, vv, w), masks = ctx.get_velocity_field()
            with eng.FlipContext(I, J, K, dx) as fresh:
                (fu, fv, fw), fmasks = fresh.velocity_advector_advect(pos[keep], vel[keep], *[a[keep] for a in aff],
                                                                      radius=radius, method=1)
            for a, b in zip(masks, fmasks):
                assert np.array_equal(a, b)
            assert close(u, fu) and close(vv, fv) and close(w, fw)
        # removing again changes nothing but the extreme-speed rule (its limit follows the new maximum)
        again, _ = ctx.remove_marker_particles(dt, 5.0, max_particles_per_cell=max(cap, 1), extreme_velocity_removal=False)
        assert again == remaining


@pytest.mark.parametrize("name", ["liquid_sdf_23x21x25_seams", "liquid_sdf_22x24x20_radius2"])
def test_liquid_sdf_golden(eng, name):
    """ParticleLevelSet::calculateSignedDistanceField on the device == the unmodified reference, bit for bit;
    resident and host-array entry points, unsorted and sorted particles."""
    meta, e = load_golden(name)
    _, src = load_golden(meta["source"])
    pos = src[meta["key"]]
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        ctx.set_particles(pos, np.zeros_like(pos))
        assert bits_equal(ctx.liquid_sdf(meta["radius"]), e["out_phi"])
        ctx.sort_particles()
        assert bits_equal(ctx.liquid_sdf(meta["radius"]), e["out_phi"])
        assert bits_equal(ctx.calculate_signed_distance_field(pos[::-1].copy(), meta["radius"]), e["out_phi"])
        ctx.declare_resident(particles=True)
        assert bits_equal(ctx.calculate_signed_distance_field(None, meta["radius"]), e["out_phi"])
        far = np.float32(3.0 * meta["dx"])
        assert bits_equal(ctx.calculate_signed_distance_field(pos[:0], meta["radius"]), np.full_like(e["out_phi"], far))


def test_liquid_sdf_vs_oracle_boundaries(eng, oracle):
    """Particles on and beyond the grid boundary, block seams and a grid that is not a multiple of the block width."""
    I, J, K, dx = 23, 31, 12, 0.013
    rng = np.random.default_rng(77)
    pos = (rng.random((40000, 3)) * [I * dx, J * dx, K * dx]).astype(np.float32)
    pos[:3000] = (rng.random((3000, 3)) * [I * dx * 1.2, J * dx * 1.2, K * dx * 1.2] - 0.1 * I * dx).astype(np.float32)
    seams = rng.integers(0, 3, (4000, 3)) * np.float32(10 * dx)
    pos[3000:7000] = (seams + rng.normal(0, 0.02 * dx, (4000, 3))).astype(np.float32)
    for radius in (0.5 * dx * np.sqrt(3.0), dx * np.sqrt(3.0), 0.3 * dx):
        want = oracle.liquid_sdf(I, J, K, dx, pos, radius)
        with eng.FlipContext(I, J, K, dx) as ctx:
            got = ctx.calculate_signed_distance_field(pos, radius)
        assert bits_equal(got, want)


SDF_VARIANT_CHECK = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import load_golden
from blender_flip_fluids_b200 import engine
from oracle import flip_oracle as fo
for name in ("liquid_sdf_23x21x25_seams", "liquid_sdf_22x24x20_radius2"):
    meta, e = load_golden(name)
    _, src = load_golden(meta["source"])
    with engine.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        got = ctx.calculate_signed_distance_field(src[meta["key"]], meta["radius"])
    assert got.tobytes() == e["out_phi"].tobytes(), name
I, J, K, dx = 23, 31, 12, 0.013
rng = np.random.default_rng(77)
pos = (rng.random((40000, 3)) * [I * dx, J * dx, K * dx]).astype(np.float32)
pos[:3000] = (rng.random((3000, 3)) * [I * dx * 1.2, J * dx * 1.2, K * dx * 1.2] - 0.1 * I * dx).astype(np.float32)
pos[3000:7000] = (rng.integers(0, 3, (4000, 3)) * np.float32(10 * dx) + rng.normal(0, 0.02 * dx, (4000, 3))).astype(np.float32)
for radius in (0.5 * dx * np.sqrt(3.0), dx * np.sqrt(3.0), 0.3 * dx):
    with engine.FlipContext(I, J, K, dx) as ctx:
        got = ctx.calculate_signed_distance_field(pos, radius)
    assert got.tobytes() == fo.liquid_sdf(I, J, K, dx, pos, radius).tobytes(), radius
meta, e = load_golden("liquid_sdf_post_24x20x22")
_, src = load_golden(meta["source"])
with engine.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
    ctx.set_solid(src[meta["solid_key"]], np.zeros(ctx.near_dims, np.uint8))
    ctx.set_particles(src[meta["key"]], np.zeros_like(src[meta["key"]]))
    assert ctx.liquid_sdf(meta["radius"]).tobytes() == e["out_phi"].tobytes()
    assert ctx.postprocess_liquid_sdf().tobytes() == e["out_phi_post"].tobytes()
print("ok")
'''


@pytest.mark.parametrize("variant", ["0", "1"])
def test_liquid_sdf_variants_and_postprocess(variant):
    """Both scatter kernels (1 = per-axis lists, the default; 0 = the first version) and the post-process kernel
    against the reference-generated fixtures, bit for bit (a fresh process per variant: the switch is read once)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e["FFB200_SDF_VARIANT"] = variant
    r = subprocess.run([sys.executable, "-c", SDF_VARIANT_CHECK, root], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("name", ["attribute_24x20x22_r1", "attribute_23x21x25_seams_r2"])
def test_attribute_transfer_golden(eng, name):
    """AttributeToGridTransfer<float> and <vmath::vec3> (attributetogridtransfer.h:52-157) at radii of 1 dx and 2 dx (the
    second on the seam-adversarial particle set) against the reference-generated fixtures: valid masks bit-exact, grids
    1e-5 (summation order), the vec3 flavour with its own normalisation."""
    meta, e = load_golden(name)
    _, src = load_golden(meta["source"])
    pos = src[meta["key"]]
    attr = (np.random.default_rng(meta["seed"]).random(len(pos)) * 10.0).astype(np.float32)
    attr3 = np.random.default_rng(meta["seed"] + 1000).random((len(pos), 3)).astype(np.float32)
    with eng.AttributeTransfer(meta["I"], meta["J"], meta["K"], meta["dx"]) as tr:
        grid, valid = tr.transfer(pos, attr, meta["radius"])
        grid3, valid3 = tr.transfer(pos, attr3, meta["radius"])
        tr.ctx.set_valid_guard(1e30, 0.0)                      # every cell through the reference-order summation
        exact, valid_e = tr.transfer(pos, attr, meta["radius"])
        exact3, _ = tr.transfer(pos, attr3, meta["radius"])
    assert np.array_equal(valid, e["out_valid"]) and np.array_equal(valid3, e["out_valid"]) and np.array_equal(valid_e, e["out_valid"])
    assert close(grid, e["out_grid"]) and close(grid3, e["out_grid3"])
    assert bits_equal(exact, e["out_grid"]) and bits_equal(exact3, e["out_grid3"])


@pytest.mark.parametrize("radius_cells", [1.0, 2.0, 3.0])
def test_attribute_transfer_vs_oracle_radii(eng, oracle, radius_cells):
    """The radii the reference uses (age / lifetime / density / colour 1 dx, whitewater proximity and viscosity solver 2 dx,
    viscosity 3 dx; fluidsimulation.h:2416-2439) on a non-dyadic grid against the oracle, scalar and vec3, normalised and not."""
    from blender_flip_fluids_b200 import scenes
    I, J, K, dx = 22, 19, 21, 0.0137
    sc = scenes.dam_break(22, dx=dx, vel="zero", dims=(I, J, K), seed=21)
    rng = np.random.default_rng(5)
    attr = (rng.random(sc.n) * 4.0 - 1.0).astype(np.float32)
    attr3 = rng.random((sc.n, 3)).astype(np.float32)
    radius = float(np.float32(radius_cells) * dx)
    og, ov = oracle.attribute_p2g(I, J, K, dx, sc.pos, attr, radius)
    og3, ov3 = oracle.attribute_p2g_vec3(I, J, K, dx, sc.pos, attr3, radius)
    with eng.AttributeTransfer(I, J, K, dx) as tr:
        g, v = tr.transfer(sc.pos, attr, radius)
        g3, v3 = tr.transfer(sc.pos, attr3, radius)
        raw, vr = tr.transfer(sc.pos, np.ones(sc.n, np.float32), radius, normalize=False)
        wsum, _ = tr.transfer(sc.pos, np.ones(sc.n, np.float32), radius, normalize=True)
    assert np.array_equal(v, ov) and np.array_equal(v3, ov3) and np.array_equal(vr, ov)
    assert close(g, og) and close(g3, og3)
    assert v.sum() > 1000
    # unnormalised transfer of a constant 1 is the weight sum itself: above 1e-6 exactly where valid, and its
    # normalised twin is 1 there
    assert (raw[ov == 1] > 1e-6).all() and np.abs(wsum[ov == 1] - 1.0).max() < 1e-5


# ---- tolerance mode (ffb200_set_precision(FFB200_PRECISION_TOLERANCE)): fp32 gathers, the north star's 1e-5 bar ----

def _tol_pos(a, b, rtol=RTOL):
    """positions: |a - b| <= 1e-5 * max(|b|, max|b|), the bar of close(), per array."""
    return close(a, b, rtol)


@pytest.mark.parametrize("name", SCENES)
def test_tolerance_mode_scene_chain(eng, name):
    """The reference-generated whole-substep fixtures (reference-built solid SDFs, dyadic and non-dyadic dx) with the
    fp32 gathers: velocities, APIC rows and advected positions within 1e-5; P2G (mode independent) unchanged."""
    meta, g = load_golden(name)
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    apic = meta["method"] == "apic"
    mac = (g["s2_u"], g["s2_v"], g["s2_w"])
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_precision(True)
        if apic:
            vel, ax, ay, az = ctx.update_marker_particle_velocities(g["s0_pos"], g["s0_vel"], mac, method=eng.APIC)
            assert close(ax, g["s3_affx"]) and close(ay, g["s3_affy"]) and close(az, g["s3_affz"])
        else:
            vel = ctx.update_marker_particle_velocities(g["s0_pos"], g["s0_vel"], mac,
                                                        saved=(g["s2_su"], g["s2_sv"], g["s2_sw"]), method=eng.FLIP,
                                                        ratio_pic_flip=meta["ratio"])
        assert close(vel, g["s3_vel"])
        ctx.tolerance_stats(reset=True)
        out = ctx.advance_marker_particles(g["s0_pos"], mac, g["s2_phi"], g["s2_near"], dt=meta["dt"], cfl=meta["cfl"])
        st = ctx.tolerance_stats()
    assert _tol_pos(out, g["s4_pos"])
    assert st["advected"] == g["s0_pos"].shape[0] and 0 <= st["advected_exact"] <= st["advected"]


def test_tolerance_mode_collision_fixture(eng):
    """The collision-heavy fixture: every particle that touches the solid must have taken the exact path (its
    projected position is a discrete outcome: anything else would be off by ~0.1 dx, 4 orders above the bar)."""
    meta, g = load_golden("advect_collide_24x20x22")
    with eng.FlipContext(meta["I"], meta["J"], meta["K"], meta["dx"]) as ctx:
        exact = ctx.advance_marker_particles(g["in_pos"], (g["in_u"], g["in_v"], g["in_w"]), g["in_phi"], g["in_near"],
                                             dt=meta["dt"], cfl=meta["cfl"])
        assert bits_equal(exact, g["out_pos"])
        ctx.set_precision(True)
        ctx.tolerance_stats(reset=True)
        out = ctx.advance_marker_particles(g["in_pos"], (g["in_u"], g["in_v"], g["in_w"]), g["in_phi"], g["in_near"],
                                           dt=meta["dt"], cfl=meta["cfl"])
        st = ctx.tolerance_stats()
        ctx.set_precision(False)
        again = ctx.advance_marker_particles(g["in_pos"], (g["in_u"], g["in_v"], g["in_w"]), g["in_phi"], g["in_near"],
                                             dt=meta["dt"], cfl=meta["cfl"])
    assert _tol_pos(out, g["out_pos"])
    assert bits_equal(again, g["out_pos"])                     # the switch goes both ways
    assert st["advected"] == g["in_pos"].shape[0]
    print("tolerance-mode fallback rate on the collision fixture:", st["advected_exact"] / max(1, st["advected"]))


@pytest.mark.parametrize("method", ["flip", "apic"])
@pytest.mark.parametrize("n,dx", [(32, 1.0 / 32), (30, 0.004 * 250 / 30), (40, 0.0123)])
def test_tolerance_mode_vs_oracle(eng, oracle, method, n, dx):
    """Dam break with a sphere obstacle on a white-noise-like P2G field (the worst case for an fp32 fraction, SURVEY hard
    part 3), dyadic and non-dyadic dx: G2P and advection in tolerance mode against the oracle at 1e-5, through the
    resident stages (k1 reuse included) and with collisions on and off."""
    from blender_flip_fluids_b200 import scenes
    apic = method == "apic"
    sc = scenes.dam_break(n, apic=apic, dx=dx, vel="random", v0=0.5, seed=11)
    I = J = K = n
    m = eng.APIC if apic else eng.FLIP
    aff = (sc.affx, sc.affy, sc.affz)
    (ou, ov, ow), _ = oracle.p2g(I, J, K, dx, sc.radius, m, sc.pos, sc.vel, *aff)
    phi, near = scenes.analytic_solid_sdf(I, J, K, dx, sphere=(0.3 * n * dx, 0.3 * n * dx, 0.5 * n * dx, 0.12 * n * dx))
    dt = 2.5 * dx / 0.5
    saved = (ou * 0.9, ov * 0.9, ow * 0.9)
    res = {}
    with eng.FlipContext(I, J, K, dx) as ctx:
        ctx.set_precision(True)
        ctx.set_solid(phi, near)
        for collide in (True, False):
            ctx.set_particles(sc.pos, sc.vel, *aff)
            ctx.set_velocity_field(ou, ov, ow)
            ctx.set_velocity_field(*saved, saved=True)
            ctx.sort_particles()
            ctx.g2p(m, 0.05)
            _, vel, ax, ay, az = ctx.get_particles(pos=False, vel=True, affine=apic)
            ctx.tolerance_stats(reset=True)
            ctx.advect(dt, 5.0, collide)
            res[collide] = (ctx.get_particles(pos=True, vel=False)[0], ctx.tolerance_stats())
    if apic:
        ovel, oax, oay, oaz = oracle.g2p_apic(I, J, K, dx, sc.pos, (ou, ov, ow))
        assert close(ax, oax) and close(ay, oay) and close(az, oaz)
    else:
        ovel = oracle.g2p_flip(I, J, K, dx, sc.pos, sc.vel, (ou, ov, ow), saved, 0.05)
    assert close(vel, ovel)
    for collide in (True, False):
        opos = oracle.advect(I, J, K, dx, sc.pos, (ou, ov, ow), phi, near, dt, 5.0, collide)
        pos1, st = res[collide]
        assert _tol_pos(pos1, opos), f"collide={collide}"
        assert st["advected"] == sc.n
        if not collide:
            assert st["advected_exact"] == 0
        else:
            assert 0 < st["advected_exact"] < sc.n            # wall/obstacle particles fall back, the interior does not
