"""CPU suite: the C oracle (oracle/flip_oracle.c) against the reference-generated golden
fixtures in tests/golden/ (tests/golden/make_golden.py). Everything is bit-exact: the oracle
repeats the reference's float/double operation order (no tolerance anywhere in this file).
"""
import numpy as np
import pytest

from conftest import bits_equal, load_golden

P2G_FIXTURES = ["p2g_flip_23x21x25_seams", "p2g_apic_23x21x25_seams", "p2g_apic_20x20x20_dyadic",
                "p2g_flip_21x20x22_radius2"]
SCENES = ["scene_flip_24x20x22_nondyadic", "scene_apic_22x24x20_dyadic"]


def _method(oracle, meta):
    return oracle.APIC if meta["method"] == "apic" else oracle.FLIP


@pytest.mark.parametrize("name", P2G_FIXTURES)
def test_p2g_stage_fixture(oracle, name):
    meta, g = load_golden(name)
    aff = [g.get("in_aff" + c) for c in "xyz"]
    (u, v, w), (vu, vv, vw) = oracle.p2g(meta["I"], meta["J"], meta["K"], meta["dx"], meta["radius"],
                                         _method(oracle, meta), g["in_pos"], g["in_vel"], *aff)
    for got, key in ((u, "out_u"), (v, "out_v"), (w, "out_w")):
        assert bits_equal(got, g[key]), key
    for got, key in ((vu, "out_validu"), (vv, "out_validv"), (vw, "out_validw")):
        assert np.array_equal(got, g[key]), key
    assert vu.sum() > 0 and vu.sum() < vu.size          # ragged: both valid and invalid faces exist


@pytest.mark.parametrize("name", SCENES)
def test_scene_chain(oracle, name):
    meta, g = load_golden(name)
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    apic = meta["method"] == "apic"
    aff = [g.get("s0_aff" + c) for c in "xyz"]
    (u, v, w), (vu, vv, vw) = oracle.p2g(I, J, K, dx, meta["radius"], _method(oracle, meta), g["s0_pos"], g["s0_vel"],
                                         *aff)
    assert bits_equal(u, g["s1_u"]) and bits_equal(v, g["s1_v"]) and bits_equal(w, g["s1_w"])
    assert np.array_equal(vu, g["s1_validu"]) and np.array_equal(vv, g["s1_validv"]) and np.array_equal(vw, g["s1_validw"])
    mac = (g["s2_u"], g["s2_v"], g["s2_w"])
    if apic:
        vel, ax, ay, az = oracle.g2p_apic(I, J, K, dx, g["s0_pos"], mac)
        assert bits_equal(ax, g["s3_affx"]) and bits_equal(ay, g["s3_affy"]) and bits_equal(az, g["s3_affz"])
    else:
        vel = oracle.g2p_flip(I, J, K, dx, g["s0_pos"], g["s0_vel"], mac, (g["s2_su"], g["s2_sv"], g["s2_sw"]),
                              meta["ratio"])
    assert bits_equal(vel, g["s3_vel"])
    out = oracle.advect(I, J, K, dx, g["s0_pos"], mac, g["s2_phi"], g["s2_near"], meta["dt"], meta["cfl"])
    assert bits_equal(out, g["s4_pos"])


def test_advect_collision_fixture(oracle):
    meta, g = load_golden("advect_collide_24x20x22")
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    mac = (g["in_u"], g["in_v"], g["in_w"])
    out = oracle.advect(I, J, K, dx, g["in_pos"], mac, g["in_phi"], g["in_near"], meta["dt"], meta["cfl"])
    assert bits_equal(out, g["out_pos"])
    free = oracle.advect(I, J, K, dx, g["in_pos"], mac, g["in_phi"], g["in_near"], meta["dt"], meta["cfl"],
                         collide=False)
    collided = (free != out).any(axis=1).sum()
    assert collided > 0.05 * len(out)                   # the fixture really exercises _resolveCollision


def test_bin_sort_properties(oracle):
    meta, g = load_golden("p2g_flip_23x21x25_seams")
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    pos = g["in_pos"].copy()
    pos[:5] = [[-1e-3, 0.01, 0.01], [0.01, (J + 1) * dx, 0.01], [0.01, 0.01, K * dx + 1], [(I + 0.5) * dx, 0.0, 0.0], [0, 0, 0]]
    cell, hkey, perm = oracle.bin_sort(I, J, K, dx, pos)
    inv = 1.0 / dx
    ci = np.floor(pos.astype(np.float64) * inv).astype(np.int64)           # grid3d.h:55-60
    ok = ((ci >= 0) & (ci < [I, J, K])).all(axis=1)
    want = np.where(ok, ci[:, 0] + I * (ci[:, 1] + J * ci[:, 2]), -1)
    assert np.array_equal(cell, want)
    assert list(ok[:5]) == [False, False, False, False, True]
    # half-cell keys refine the cell index exactly (apron of 4 half-cells per side)
    A, HX, HY, HZ = 4, 2 * I + 8, 2 * J + 8, 2 * K + 8
    hk = hkey[ok].astype(np.int64)
    hi, hj, hk2 = hk % HX - A, (hk // HX) % HY - A, hk // (HX * HY) - A
    assert np.array_equal((hi >> 1) + I * ((hj >> 1) + J * (hk2 >> 1)), cell[ok])
    assert hkey[2] == HX * HY * HZ                      # far outside: sentinel bin
    assert hkey[0] < HX * HY * HZ                       # just outside: still binned in the apron
    # stable order: keys ascending, ties by ascending particle index
    ks = hkey[perm]
    assert (np.diff(ks.astype(np.int64)) >= 0).all()
    same = np.diff(ks.astype(np.int64)) == 0
    assert (np.diff(perm.astype(np.int64))[same] > 0).all()
    assert np.array_equal(np.sort(perm), np.arange(len(pos)))


EXTRAPOLATE = [("extrapolate_flip_24x20x22", "scene_flip_24x20x22_nondyadic"),
               ("extrapolate_apic_22x24x20", "scene_apic_22x24x20_dyadic")]


@pytest.mark.parametrize("name,scene", EXTRAPOLATE)
def test_extrapolate_fixture(oracle, name, scene):
    """GridUtils::extrapolateGrid: the reference's output (identical for 1, 3 and 16 reference threads,
    checked by make_golden.py) on its own P2G output + valid masks."""
    meta, e = load_golden(name)
    _, g = load_golden(scene)
    assert meta["layers"] == oracle.extrapolation_layers(5.0) == 12
    for c in "uvw":
        out = oracle.extrapolate(g["s1_" + c], g["s1_valid" + c], meta["layers"])
        assert bits_equal(out, e["out_" + c]), c
        assert (out != g["s1_" + c]).sum() > 0               # something was extrapolated
        assert bits_equal(out[g["s1_valid" + c] == 1], g["s1_" + c][g["s1_valid" + c] == 1])   # valid faces untouched


def test_extrapolate_edge_cases(oracle):
    rng = np.random.default_rng(5)
    grid = rng.standard_normal((6, 7, 8)).astype(np.float32)
    none = np.zeros(grid.shape, np.uint8)
    assert bits_equal(oracle.extrapolate(grid, none, 12), grid)            # nothing valid: nothing changes
    full = np.ones(grid.shape, np.uint8)
    assert bits_equal(oracle.extrapolate(grid, full, 12), grid)            # everything valid: nothing changes
    one = none.copy()
    one[3, 3, 4] = 1
    assert bits_equal(oracle.extrapolate(grid, one, 0), grid)              # zero layers
    out = oracle.extrapolate(grid, one, 1)
    # first layer: the six neighbours take the mean of their DONE neighbours = the seed (border cells
    # are DONE from the start and join the mean where they touch)
    assert out[3, 3, 5] == grid[3, 3, 4] and out[3, 3, 3] == grid[3, 3, 4]
    changed = np.argwhere(out != grid)
    assert len(changed) <= 6


def test_remove_particles_open_boundaries_fixture(oracle):
    """Same with three open domain sides (fluidsimulation.cpp:7780-7823), and the caller-evaluated pre-removal mask."""
    meta, g = load_golden("remove_open_24x20x22")
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    args = (I, J, K, dx, g["in_pos"], g["in_vel"], g["in_phi"], meta["dt"], meta["cfl"])
    removed, extreme = oracle.remove_particles(*args, open_bounds=g["in_bounds"])
    keep = removed == 0
    assert int(keep.sum()) == meta["survivors"] and extreme == meta["extreme"]
    assert bits_equal(g["in_pos"][keep], g["out_pos"]) and bits_equal(g["in_vel"][keep], g["out_vel"])
    closed, _ = oracle.remove_particles(*args)
    assert (closed == 0).sum() > keep.sum() + 100                     # the open sides really remove particles
    # a pre-removed particle is dropped without taking a slot of its cell: same as deleting it beforehand
    pre = (np.arange(len(removed)) % 7 == 0).astype(np.uint8)
    with_pre, _ = oracle.remove_particles(*args, max_per_cell=3, extreme_removal=False, pre_removed=pre)
    sub = pre == 0
    without, _ = oracle.remove_particles(I, J, K, dx, g["in_pos"][sub], g["in_vel"][sub], g["in_phi"], meta["dt"], meta["cfl"],
                                         max_per_cell=3, extreme_removal=False)
    assert (with_pre[pre == 1] == 1).all() and np.array_equal(with_pre[sub], without)


def test_remove_particles_fixture(oracle):
    """FluidSimulation::_removeMarkerParticles (fluidsimulation.cpp:7723-7851): survivors of the unmodified
    reference, in order."""
    meta, g = load_golden("remove_24x20x22")
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    removed, extreme = oracle.remove_particles(I, J, K, dx, g["in_pos"], g["in_vel"], g["in_phi"], meta["dt"], meta["cfl"])
    keep = removed == 0
    assert int(keep.sum()) == meta["survivors"] < meta["particles"]
    assert extreme == meta["extreme"] > 0
    assert bits_equal(g["in_pos"][keep], g["out_pos"]) and bits_equal(g["in_vel"][keep], g["out_vel"])
    # each of the three rules removes something in this fixture: solid, crowded cell, extreme speed
    no_extreme, n0 = oracle.remove_particles(I, J, K, dx, g["in_pos"], g["in_vel"], g["in_phi"], meta["dt"], meta["cfl"],
                                             extreme_removal=False)
    assert n0 == 0 and 0 < no_extreme.sum() < removed.sum()
    no_cap, _ = oracle.remove_particles(I, J, K, dx, g["in_pos"], g["in_vel"], g["in_phi"], meta["dt"], meta["cfl"],
                                        max_per_cell=1 << 30, extreme_removal=False)
    assert 0 < no_cap.sum() < no_extreme.sum()
    # per-cell cap keeps the first 250 of a cell in index order
    ci = np.floor(g["in_pos"].astype(np.float64) / dx).astype(np.int64)
    crowded = (ci == [9, 8, 10]).all(axis=1) & (no_cap == 0)
    assert crowded.sum() > 250 and (no_extreme[crowded] == 0).sum() == 250
    assert (no_extreme[np.flatnonzero(crowded)[:250]] == 0).all()


LIQUID_SDF = ["liquid_sdf_23x21x25_seams", "liquid_sdf_22x24x20_radius2", "liquid_sdf_post_24x20x22"]


@pytest.mark.parametrize("name", LIQUID_SDF)
def test_liquid_sdf_fixture(oracle, name):
    """ParticleLevelSet::calculateSignedDistanceField (particlelevelset.cpp:161-168, 335-668): the reference's cell-centred
    liquid SDF (identical for 1, 3 and 16 reference threads) on fixture positions; groundwork for SURVEY §8f row f3."""
    meta, e = load_golden(name)
    _, src = load_golden(meta["source"])
    I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
    pos = src[meta["key"]]
    phi = oracle.liquid_sdf(I, J, K, dx, pos, meta["radius"])
    assert bits_equal(phi, e["out_phi"])
    far = np.float32(3.0 * dx)
    assert phi.max() == far and (phi < 0).sum() > 1000 and (phi == far).sum() > 1000
    # a minimum over particles: any particle order gives the same field
    perm = np.random.default_rng(2).permutation(len(pos))
    assert bits_equal(oracle.liquid_sdf(I, J, K, dx, pos[perm], meta["radius"]), phi)
    assert bits_equal(oracle.liquid_sdf(I, J, K, dx, pos[:0], meta["radius"]), np.full_like(phi, far))
    if meta.get("solid_key"):
        # ParticleLevelSet::postProcessSignedDistanceField (particlelevelset.cpp:170-195) against the fixture's solid SDF
        post = oracle.liquid_sdf_postprocess(I, J, K, dx, phi, src[meta["solid_key"]])
        assert bits_equal(post, e["out_phi_post"])
        assert (post == np.float32(-0.5 * dx)).sum() > 500 and np.abs(post).min() >= np.float32(0.005 * dx)


def test_liquid_sdf_axes_decomposition(oracle):
    """The per-axis evaluation with the squared-distance pre-filter (what the device's opt-in k_sdf_scatter_axes does) is
    bit-identical to the pinned restatement, on the fixtures and on a boundary / block-seam stress cloud."""
    for name in LIQUID_SDF:
        meta, e = load_golden(name)
        _, src = load_golden(meta["source"])
        phi, skipped = oracle.liquid_sdf_axes(meta["I"], meta["J"], meta["K"], meta["dx"], src[meta["key"]], meta["radius"])
        assert bits_equal(phi, e["out_phi"]) and skipped > 0
    I, J, K, dx = 23, 31, 12, 0.013
    rng = np.random.default_rng(77)
    pos = (rng.random((20000, 3)) * [I * dx, J * dx, K * dx]).astype(np.float32)
    pos[:2000] = (rng.random((2000, 3)) * [I * dx * 1.2, J * dx * 1.2, K * dx * 1.2] - 0.1 * I * dx).astype(np.float32)
    pos[2000:5000] = (rng.integers(0, 3, (3000, 3)) * np.float32(10 * dx) + rng.normal(0, 0.02 * dx, (3000, 3))).astype(np.float32)
    for radius in (0.5 * dx * np.sqrt(3.0), dx * np.sqrt(3.0), 0.3 * dx, 2.2 * dx):
        phi, _ = oracle.liquid_sdf_axes(I, J, K, dx, pos, radius)
        assert bits_equal(phi, oracle.liquid_sdf(I, J, K, dx, pos, radius)), radius


def test_attribute_p2g_is_a_shifted_u_transfer(oracle):
    """The identity behind engine.AttributeTransfer: the attribute transfer onto I x J x K equals the U-direction FLIP
    transfer on (I-1) x J x K with x shifted by float(dx/2) and the attribute in the x velocity -- bit for bit."""
    for name in ATTRIBUTE:
        meta, e = load_golden(name)
        _, src = load_golden(meta["source"])
        I, J, K, dx = meta["I"], meta["J"], meta["K"], meta["dx"]
        pos = src[meta["key"]]
        attr = (np.random.default_rng(meta["seed"]).random(len(pos)) * 10.0).astype(np.float32)
        shifted = pos.copy()
        shifted[:, 0] = pos[:, 0] - np.float32(0.5 * dx)
        vel = np.zeros_like(pos)
        vel[:, 0] = attr
        (u, _, _), (vu, _, _) = oracle.p2g(I - 1, J, K, dx, meta["radius"], 0, shifted, vel)
        assert bits_equal(u, e["out_grid"]) and np.array_equal(vu, e["out_valid"])


ATTRIBUTE = ["attribute_23x21x25_seams_r2", "attribute_24x20x22_r1"]


@pytest.mark.parametrize("name", ATTRIBUTE)
def test_attribute_p2g_fixture(oracle, name):
    """AttributeToGridTransfer<float>::transfer (attributetogridtransfer.h:52-157): the reference's cell-centred attribute
    grid + valid mask at radii of 2 dx and 1 dx (identical for 1 and 16 reference threads); SURVEY §8f row f4."""
    meta, e = load_golden(name)
    _, src = load_golden(meta["source"])
    pos = src[meta["key"]]
    attr = (np.random.default_rng(meta["seed"]).random(len(pos)) * 10.0).astype(np.float32)
    grid, valid = oracle.attribute_p2g(meta["I"], meta["J"], meta["K"], meta["dx"], pos, attr, meta["radius"])
    assert np.array_equal(valid, e["out_valid"]) and bits_equal(grid, e["out_grid"])
    assert valid.sum() > 1500 and np.abs(grid[valid == 0]).max() < 1e-4      # weight <= 1e-6: unnormalised leftovers, not valid
    # AttributeToGridTransfer<vmath::vec3> (colour): same sums per channel, but normalised by vec3 /= float (x * float(1/w))
    attr3 = np.random.default_rng(meta["seed"] + 1000).random((len(pos), 3)).astype(np.float32)
    grid3, valid3 = oracle.attribute_p2g_vec3(meta["I"], meta["J"], meta["K"], meta["dx"], pos, attr3, meta["radius"])
    assert np.array_equal(valid3, valid) and bits_equal(grid3, e["out_grid3"])
    # a constant attribute comes back as that constant wherever the grid is valid (normalised weights)
    ones, v1 = oracle.attribute_p2g(meta["I"], meta["J"], meta["K"], meta["dx"], pos, np.full(len(pos), 3.0, np.float32), meta["radius"])
    assert np.array_equal(v1, valid) and np.abs(ones[valid == 1] - 3.0).max() < 1e-5
