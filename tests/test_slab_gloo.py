"""CPU suite over gloo: the z-slab driver (ghost particles, face halos, migration) against an
undecomposed oracle run, bit-for-bit, with particles crossing the slab boundaries. Two code paths:
the generic plumbing (world size 2) and the device-plumbing protocol of step_fast -- one merged
migrant + ghost exchange per substep, per-face buffer sizes agreed through the headers -- with the
CUDA kernels restated in numpy (world sizes 2 and 3: a middle rank has two faces). Spawns the
processes on 127.0.0.1."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _scene(apic, K=28):
    from blender_flip_fluids_b200 import scenes
    I, J, dx = 12, 10, 0.05
    sc = scenes.dam_break(12, apic=apic, dx=dx, dims=(I, J, K), vel="random", v0=0.6, seed=21)
    kz = np.floor(sc.pos[:, 2].astype(np.float64) / dx).astype(np.int64)
    sc.vel[:, 2] += np.where(kz % 14 >= 7, 0.9, -0.9).astype(np.float32)   # drive particles across the slab faces (k = 14, 28)
    phi, near = scenes.analytic_solid_sdf(I, J, K, dx)
    return I, J, K, dx, sc, phi, near


def _streams(sc, apic, sel):
    cols = [sc.pos[sel, 0], sc.pos[sel, 1], sc.pos[sel, 2], sc.vel[sel, 0], sc.vel[sel, 1], sc.vel[sel, 2]]
    if apic:
        for a in (sc.affx, sc.affy, sc.affz):
            cols += [a[sel, 0], a[sel, 1], a[sel, 2]]
    return [torch.from_numpy(np.ascontiguousarray(c)) for c in cols]


def _worker(rank, world, port, apic, out, fast=False, K=28, steps=2):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from blender_flip_fluids_b200 import slab
        from cpu_slab_backend import CpuOracleBackend, CpuOracleFastBackend
        I, J, K, dx, sc, phi, near = _scene(apic, K)
        kb, ke = slab.slab_range(K, world, rank)
        be = (CpuOracleFastBackend if fast else CpuOracleBackend)(I, J, K, dx, kb, ke, 7, apic)
        be.set_solid(phi, near)
        sim = slab.SlabSimulation(I, J, K, dx, rank, world, be, halo=7, ghost=int(os.environ.get("FFB200_TEST_GHOST", "2")))
        kz = np.floor(sc.pos[:, 2].astype(np.float64) * (1.0 / dx)).astype(np.int64)
        sel = np.nonzero((kz >= kb) & (kz < ke))[0]
        sim.set_particles(_streams(sc, apic, sel), torch.from_numpy(sel.astype(np.int32)))
        if os.environ.get("FFB200_TEST_TINY_CAPS") == "1":
            sim.force_initial_caps = (8, 16, 8, 16)          # every merged exchange overflows and is repeated once
        dt = 1.5 * dx / 1.5
        moved = 0
        if fast:
            sim.load_resident()
        for _ in range(steps):
            before = set(sim.ids.tolist())
            if fast:
                sim.step_fast(sc.radius, 0.05, dt)
                sim.sync_from_backend()
            else:
                sim.step(sc.radius, 0.05, dt)
            moved += len(set(sim.ids.tolist()) - before)
        allp, ids = sim.gather_particles()
        if rank == 0:
            np.savez(out, streams=allp.numpy(), ids=ids.numpy(), moved=moved, exchanged=sim.exchanged_bytes,
                     overflows=getattr(sim, "overflows", 0))
    finally:
        dist.destroy_process_group()


def _run_and_compare(tmp_path, oracle, apic, world, fast, K, steps):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "slab.npz")
    mp.spawn(_worker, args=(world, port, apic, out, fast, K, steps), nprocs=world, join=True)
    got = np.load(out)
    I, J, K, dx, sc, phi, near = _scene(apic, K)
    pos, vel = sc.pos.copy(), sc.vel.copy()
    aff = [sc.affx, sc.affy, sc.affz] if apic else [None] * 3
    dt = 1.5 * dx / 1.5
    for _ in range(steps):                                  # undecomposed: P2G -> save -> G2P -> advect
        (u, v, w), _ = oracle.p2g(I, J, K, dx, sc.radius, oracle.APIC if apic else oracle.FLIP, pos, vel, *aff)
        if apic:
            vel, ax, ay, az = oracle.g2p_apic(I, J, K, dx, pos, (u, v, w))
            aff = [ax, ay, az]
        else:
            vel = oracle.g2p_flip(I, J, K, dx, pos, vel, (u, v, w), (u, v, w), 0.05)
        pos = oracle.advect(I, J, K, dx, pos, (u, v, w), phi, near, dt, 5.0, True)
    want = [pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2]]
    if apic:
        for a in aff:
            want += [a[:, 0], a[:, 1], a[:, 2]]
    assert np.array_equal(got["ids"], np.arange(sc.n))
    assert int(got["moved"]) > 0, "the scene must exercise migration"
    for q, wq in enumerate(want):
        assert got["streams"][q].tobytes() == np.ascontiguousarray(wq).tobytes(), f"stream {q} differs"


@pytest.mark.parametrize("apic", [False, True])
def test_slab_two_ranks_match_single_domain(tmp_path, apic, oracle):
    _run_and_compare(tmp_path, oracle, apic, world=2, fast=False, K=28, steps=2)


@pytest.mark.parametrize("world,apic,ghost", [(2, True, 2), (3, False, 2), (3, True, 2), (3, True, 1), (2, False, 1)])
def test_slab_fast_protocol_matches_single_domain(tmp_path, world, apic, ghost, oracle, monkeypatch):
    """step_fast's exchange protocol (merged migrant + ghost exchange, ghosts kept by the sender,
    per-face capacities from the headers) with the kernels restated in numpy: three substeps, three
    ranks (a middle rank exchanges on two faces), bit-identical to the undecomposed run. One ghost particle layer is
    enough for the default kernel radius of 0.866 dx (what bench.py uses); two cover the doubled radius."""
    monkeypatch.setenv("FFB200_TEST_GHOST", str(ghost))
    _run_and_compare(tmp_path, oracle, apic, world=world, fast=True, K=14 * world, steps=3)


def test_slab_exchange_repeats_overflowing_sections(tmp_path, oracle, monkeypatch):
    """Sections far too small for the migrants and ghost copies of a face: the headers carry the true counts, both
    ranks of the face repeat the exchange on that face alone with resized buffers (the marking is repeatable: kept
    migrants get their ghost bit in route_end), and the run is still bit-identical to the undecomposed one --
    nothing is lost, nothing stays behind with its sender."""
    monkeypatch.setenv("FFB200_TEST_TINY_CAPS", "1")
    monkeypatch.setenv("FFB200_TEST_GHOST", "1")
    _run_and_compare(tmp_path, oracle, True, world=3, fast=True, K=42, steps=3)
    assert int(np.load(str(tmp_path / "slab.npz"))["overflows"]) >= 3
