#!/usr/bin/env python
"""bench.py -- FLIP substep particle-updates/s (P2G + G2P + advect) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 512]

One "step" is one pass of the hot path over the scene's particles: cell binning + sort, P2G (U, V, W),
velocity-field save, G2P, RK3 advection with collision.

WORKLOAD (every N, strong scaling): BASELINE.json configs[3] -- dam break 512 x 512 x 512, APIC transfer,
8 particles per cell (330 341 088 particles), z-slab sharded across the N ranks (N = 1: one context holds the
whole scene, ~133 GB of the 180 GB). Particles come from a counter-based hash of (seed, global particle id)
(blender_flip_fluids_b200/scenes.py), so every decomposition -- and the CPU reference arm -- sees
bit-identical particles. The batch EVOLVES: every step sorts the particles the previous step advected, ranks
migrate particles for real. Where the reference's CPU pressure projection would hand back a field, the MAC
field is overwritten with an analytic divergence-free field (two superposed Taylor-Green vortices, tangential at
the walls of the inner box the fluid lives in, with motion across z), so the set circulates -- and crosses slab
faces -- instead of compressing or piling up against the solid. Same launch mode (eager) at every N.

`value`    device-resident throughput: particles and grids live in HBM, K steps timed with CUDA events on the
           launching stream between barriers, max over ranks.
`e2e`      the same step through the reference-facing host-buffer entry points of the C ABI with the particles
           RESIDENT (the protocol the libffengine interposer runs): per step the P2G's faces + valid masks go to
           pinned host memory, the projected field comes back from pinned host memory, the CFL speed is read back.
`roofline` whole-substep algorithmic bytes / step time against MEASURED_PEAKS.json, with the
           per-stage figures beside it.
`checksum` order-independent hash of the owned particles (global id, position and velocity bits) after the K+W
           steps and of the P2G field of the final state: equal across N iff the slab path reproduces the single-GPU
           bits.
`secondary` (N = 1) BASELINE.json configs[1]: dam break 128^3 APIC (4 637 952 particles): evolving and fixed batch,
           eager launches and CUDA-graph replay (the round-1 headline configuration, kept for comparison).
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference engine (oracle/_ref/ref_harness, built from
           /root/reference) on this box's host cores, all threads, on a bounded sample of the same scene (a slab of
           16 interior cell planes; the reference at 330 M particles would need minutes per step).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FLIP substep particle-updates/sec (P2G+G2P+advect)"
UNIT = "particle-updates/s"
GRID_PRIMARY = 512
GRID_SECONDARY = 128
METHOD = "apic"
PPC = 8
V0 = 0.5                      # |v| component bound of the synthetic velocities and the amplitude of the analytic field
SEED = 1234
RATIO = 0.05
HALO = 7
GHOST = 1                     # ghost particle layers: ceil(radius / dx) -- the P2G kernel reaches 0.866 dx, so 1 (2 for the doubled
                              # radius of the smooth surface-tension kernel); bit-identity vs the undecomposed run: tests/test_slab_gloo.py
# N > 1: neighbour exchanges overlapped with the interior particles' G2P + advection. It pays once the slabs are thin (the
# windowed launches cost a few % on thick slabs: 32.5 vs 30.8 ms at N = 2, 9.55 vs 10.14 ms at N = 8, 512^3); "auto" = slabs of
# at most 128 planes
OVERLAP_ENV = os.environ.get("FFB200_BENCH_OVERLAP", "auto")
SAMPLE_PLANES = 16            # fluid cell planes of the CPU reference sample
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def algorithmic_bytes(method: str, ppc: float):
    """compulsory bytes per particle-update, per stage."""
    if method == "apic":
        return {"p2g": 60 + 15 / ppc, "g2p": 60 + 12 / ppc, "advect": 24 + 16 / ppc}
    return {"p2g": 24 + 15 / ppc, "g2p": 36 + 24 / ppc, "advect": 24 + 16 / ppc}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_string(n, total):
    return (f"dam break {n}x{n}x{n} (BASELINE configs[{3 if n == 512 else 1}]), {METHOD.upper()} transfer, RK3 advection + "
            f"collision, ppc {PPC}, {total} particles, hashed generator seed {SEED}")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
def reference_arm(n_grid: int, steps: int, warmup: int, threads: int = 0):
    """Time the UNMODIFIED reference (oracle/_ref/ref_harness) on host cores: VelocityAdvector::advect,
    _updateMarkerParticleVelocitiesThread, _advanceMarkerParticlesThread on a bounded sample of the bench scene:
    SAMPLE_PLANES interior cell planes of the n^3 dam break (same x/y extent, same hashed particles, same analytic
    field; the x/y walls are solid as in the scene, the z neighbours are fluid, so the collision gate fires at the
    scene's rate), or the whole scene when it has no more planes than that. Returns (value, info)."""
    if not os.path.exists(HARNESS):
        raise RuntimeError("oracle/_ref/ref_harness missing: run __graft_entry__.build() where /root/reference exists")
    from blender_flip_fluids_b200 import scenes
    n, dx = n_grid, 1.0 / n_grid
    e = scenes.dam_break_extent(n, n, n)
    whole = (e[5] - e[4]) <= 4 * SAMPLE_PLANES
    if whole:
        K, k0, k1, shift = n, e[4], e[5], 0
    else:
        K = SAMPLE_PLANES + 6
        k0 = (n - SAMPLE_PLANES) // 2
        k1, shift = k0 + SAMPLE_PLANES, k0 - 3                       # scene plane k -> sample plane k - shift
    streams, _ = scenes.dam_break_planes(n, n, n, dx, k0, k1, apic=(METHOD == "apic"), v0=V0, seed=SEED, xp=np)
    pos = np.stack(streams[0:3], axis=1)
    pos[:, 2] = (pos[:, 2].astype(np.float64) - shift * dx).astype(np.float32)
    vel = np.stack(streams[3:6], axis=1)
    aff = [np.stack(streams[6 + 3 * q:9 + 3 * q], axis=1) for q in range(3)] if METHOD == "apic" else None
    # analytic field (benchscene.taylor_green_field restated with numpy: the same doubles, narrowed once), on the sample's
    # planes in SCENE coordinates
    A = B = 0.5 * V0
    kz0 = 0 if whole else k0 - 3
    a_in, L_in = 3.0 * dx, dx * n - 6.0 * dx
    inner = lambda t: np.clip((t - a_in) / L_in, 0.0, 1.0)
    xf, xc = inner(np.arange(n + 1) * dx), inner((np.arange(n) + 0.5) * dx)
    yf, yc = inner(np.arange(n + 1) * dx), inner((np.arange(n) + 0.5) * dx)
    zc, zf = inner((np.arange(kz0, kz0 + K) + 0.5) * dx), inner(np.arange(kz0, kz0 + K + 1) * dx)
    u2 = (A * np.sin(math.pi * xf)[None, :] * np.cos(math.pi * yc)[:, None]).astype(np.float32)
    mac = [np.ascontiguousarray(np.broadcast_to(u2[None], (K, n, n + 1))),
           np.ascontiguousarray(np.broadcast_to((-A * np.cos(math.pi * xc)[None, None, :] * np.sin(math.pi * yf)[None, :, None] +
                                                 B * np.sin(math.pi * yf)[None, :, None] * np.cos(math.pi * zc)[:, None, None]).astype(np.float32),
                                                (K, n + 1, n))),
           np.ascontiguousarray(np.broadcast_to((-B * np.cos(math.pi * yc)[None, :, None] * np.sin(math.pi * zf)[:, None, None]).astype(np.float32),
                                                (K + 1, n, n)))]
    if whole:
        phi, near = scenes.analytic_solid_sdf(n, n, n, dx)
    else:
        # x / y walls of the scene; the sample's z faces are interior planes of the scene (far from its z walls)
        inset = 0.5 * (3.0 * dx + 1e-4)
        y, x = np.meshgrid(np.arange(n + 1) * dx, np.arange(n + 1) * dx, indexing="ij")
        d2 = np.minimum.reduce([x - inset, n * dx - inset - x, y - inset, n * dx - inset - y])
        zdist = np.minimum(np.arange(k0 - 3, k0 - 3 + K + 1) * dx - inset, n * dx - inset - np.arange(k0 - 3, k0 - 3 + K + 1) * dx)
        phi = np.minimum(d2[None], zdist[:, None, None]).astype(np.float32)
        gi, gk = math.ceil(n / 3), math.ceil(K / 3)
        near = np.zeros((gk, gi, gi), np.uint8)
        band = np.abs(phi[:K, :n, :n]) < np.float32(3.0 * dx)
        kk, jj, ii = np.nonzero(band)
        near[kk // 3, jj // 3, ii // 3] = 1
        for _ in range(2):
            g = near.copy()
            g[:, 1:, :] |= near[:, :-1, :]; g[:, :-1, :] |= near[:, 1:, :]
            g[:, :, 1:] |= near[:, :, :-1]; g[:, :, :-1] |= near[:, :, 1:]
            g[1:] |= near[:-1]; g[:-1] |= near[1:]
            near = g
    d = tempfile.mkdtemp(prefix="ffb200_ref_")
    try:
        arrs = dict(pos=pos, vel=vel, u=mac[0], v=mac[1], w=mac[2], phi=phi, near=near)
        if aff:
            arrs.update(affx=aff[0], affy=aff[1], affz=aff[2])
        for k, a in arrs.items():
            np.save(os.path.join(d, f"in_{k}.npy"), a)
        reps = steps + warmup
        common = [f"I={n}", f"J={n}", f"K={K}", f"dx={dx!r}", f"method={METHOD}", f"reps={reps}", f"threads={threads}"]
        dt = 1.0 * dx / V0

        def run(mode, *extra):
            r = subprocess.run([HARNESS, mode, d] + common + list(extra), capture_output=True, text=True, check=True)
            return json.loads(r.stdout.strip().splitlines()[-1])

        a = run("p2g")
        b = run("g2p", f"ratio={RATIO}")
        c = run("advect", f"dt={dt!r}", "cfl=5")
    finally:
        shutil.rmtree(d, ignore_errors=True)
    per_step = [x + y + z for x, y, z in zip(a["times"], b["times"], c["times"])][warmup:]
    total = sum(per_step)
    m = int(pos.shape[0])
    full = scenes.dam_break_count(n, n, n, PPC)
    info = {"particles": m, "threads": a["threads"], "t_p2g": statistics.mean(a["times"][warmup:]),
            "t_g2p": statistics.mean(b["times"][warmup:]), "t_advect": statistics.mean(c["times"][warmup:]),
            "ms_per_step": 1e3 * total / len(per_step), "same_config": True, "scene_particles": full,
            "sample": (f"the whole scene ({m} particles)" if whole else
                       f"{SAMPLE_PLANES} interior cell planes k={k0}..{k1 - 1} of the scene ({m} of {full} particles, same hashed particles "
                       f"and analytic field, x/y walls solid, z neighbours fluid); the rate is per particle, i.e. extrapolated to the "
                       f"full scene") + "; stage-level VelocityAdvector::advect + _updateMarkerParticleVelocitiesThread + "
                      "_advanceMarkerParticlesThread (no extrapolation/removal) of the unmodified reference via oracle/_ref/ref_harness"}
    return m * len(per_step) / total, info


def host_cpu():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return model, os.cpu_count()


# --------------------------------------------------------------------------------------------------
class Runner:
    """One rank's share of the n^3 hashed dam break, resident, with the step in its modes."""

    def __init__(self, n, rank, world, local_rank, stream):
        import torch
        from blender_flip_fluids_b200 import benchscene, engine, slab
        self.torch, self.engine, self.bs = torch, engine, benchscene
        self.n, self.rank, self.world = n, rank, world
        self.dx = 1.0 / n
        self.dev = torch.device("cuda", local_rank)
        self.stream = stream
        self.m = engine.APIC if METHOD == "apic" else engine.FLIP
        self.apic = METHOD == "apic"
        self.radius = 0.5 * self.dx * math.sqrt(3.0)
        self.dt = self.dx / V0            # ~1 cell per substep at the velocity bound (CFL limit is 5)
        self.kb, self.ke = slab.slab_range(n, world, rank)
        self.overlap = world > 1 and ((n // world) <= 128 if OVERLAP_ENV == "auto" else OVERLAP_ENV != "0")
        if world > 1:
            self.backend = slab.GpuBackend(n, n, n, self.dx, self.kb, self.ke, HALO, local_rank, self.apic)
            self.ctx = self.backend.ctx
            self.sim = slab.SlabSimulation(n, n, n, self.dx, rank, world, self.backend, halo=HALO, ghost=GHOST)
        else:
            self.ctx = engine.FlipContext(n, n, n, self.dx, device=local_rank)
            self.ctx.set_stream(stream.cuda_stream)
            self.sim = None
        benchscene.set_wall_solid(self.ctx, n, n, n, self.dx, self.dev)
        self.n_local = benchscene.fill_dam_break(self.ctx, n, n, n, self.dx, self.kb, self.ke, self.apic, V0, SEED, self.dev,
                                                 headroom=1.15 if world == 1 else 1.35)
        if self.sim is not None:
            self.sim.adopt_resident(self.n_local)
        b = self.ctx.device_buffers()
        self.tg = benchscene.taylor_green_field(n, n, n, self.dx, b.kbase, b.kloc, V0, self.dev)
        self.fv = benchscene.field_views(self.ctx, self.dev)
        self.fixed = False

    def set_fixed(self, on):
        self.fixed = bool(on)
        self.ctx.set_fixed_batch(self.fixed)

    def step(self):
        if self.sim is not None:
            self.sim.step_fast(self.radius, RATIO, self.dt, apply_migration=not self.fixed,
                               projected_field=None if self.fixed else self.tg, overlap=self.overlap)
            return
        c = self.ctx
        c.p2g(self.radius, self.m)            # bins + sort + seam words + U, V, W transfers
        c.save_velocity_field()               # _saveVelocityField; the CPU pressure solve would sit here ...
        if not self.fixed:
            for a, t in zip(self.fv, self.tg):    # ... and hand back a divergence-free field
                a.copy_(t.view(-1))
        c.g2p(self.m, RATIO)
        c.advect(self.dt, 5.0, True)

    def time_steps(self, steps, barrier, replay=None):
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(self.stream)
        if replay is not None:
            for _ in range(steps // 2):
                replay()
        else:
            for _ in range(steps):
                self.step()
        ev1.record(self.stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        return ms / (2 * (steps // 2) if replay is not None else steps)

    def capture(self):
        """Two consecutive substeps (the sort flips the double buffer: two return it to where it started) in one
        CUDA graph. Single GPU only (no host decision in the step)."""
        torch = self.torch
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=self.stream):
            self.step()
            self.step()
        torch.cuda.set_stream(self.stream)
        graph.replay()
        torch.cuda.synchronize()
        return graph

    def stage_times(self, reps=3):
        stage = {"sort_ms": 0.0, "p2g_prep_ms": 0.0, "p2g_ms": 0.0, "g2p_ms": 0.0, "advect_ms": 0.0}
        launches = 0
        for _ in range(reps):
            self.step()
            t = self.ctx.timing()
            for k in stage:
                stage[k] += t[k] / reps
            launches = sum(t[k] for k in ("sort_launches", "p2g_prep_launches", "p2g_launches", "g2p_launches", "advect_launches"))
        return stage, launches

    def checksums(self):
        """(owned particle count, particle checksum, P2G field checksum of the current state), this rank's share."""
        cnt, ph = self.bs.particle_checksum(self.ctx, self.apic, self.dev)
        # the transfer of the current particles onto the faces this rank owns (ghost copies are in place after a step)
        if self.sim is not None:
            if not getattr(self.sim, "_ghosts_ready", False):
                return cnt, ph, None
            self.backend.p2g(self.radius)
        else:
            self.ctx.p2g(self.radius, self.m)
        fh = self.bs.field_checksum(self.ctx, self.dev, self.kb, self.ke, top_w=(self.rank == self.world - 1))
        return cnt, ph, fh

    def close(self):
        self.sim = None
        self.fv = self.tg = None
        self.ctx.close()
        self.torch.cuda.empty_cache()


def roofline_block(n_local, ms_per_step, stage, peak, peak_src, traffic):
    balg = algorithmic_bytes(METHOD, PPC)
    total_b = sum(balg.values()) * n_local
    ach = total_b / (ms_per_step * 1e-3) / 1e9

    def st(key, ms):
        a = balg[key] * n_local / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"algorithmic_bytes": balg[key] * n_local, "ms": ms, "achieved": a, "frac": a / peak}

    return {"kernel": "whole substep: sort + P2G (k_p2g_* x3 directions) + G2P + advect, per GPU",
            "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": total_b, "launch_ms": ms_per_step,
            "algorithmic_bytes_per_particle": sum(balg.values()),
            "stages": {"p2g (3 directions on 3 streams, incl. cell lists and node gather)": st("p2g", stage["p2g_ms"]),
                       "g2p (one kernel)": st("g2p", stage["g2p_ms"]), "advect (one kernel)": st("advect", stage["advect_ms"]),
                       "sort + membership (no algorithmic bytes)": {"ms": stage["sort_ms"] + stage["p2g_prep_ms"]}},
            "stage_ms": stage}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=GRID_PRIMARY)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-checksum", action="store_true")
    ap.add_argument("--no-tolerance", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    n = args.grid

    # stdout carries exactly one JSON line: everything libraries write to fd 1 meanwhile (NCCL prints its
    # version banner there) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    from blender_flip_fluids_b200 import scenes
    n_scene = scenes.dam_break_count(n, n, n, PPC)
    base_config = {"workload": workload_string(n, n_scene), "dt": (1.0 / n) / V0, "pic_flip_ratio": RATIO}

    if args.impl == "reference":
        if rank != 0:
            return 0
        model, cores = host_cpu()
        # every step is ~2 s of CPU work on the 16-plane sample: K steps are timed as asked, up to 40 (a run of minutes)
        ref_steps = min(steps, 40)
        value, info = reference_arm(n, ref_steps, min(warmup, 2))
        cfg = dict(base_config, cpu_model=model, timed_reps=ref_steps)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": ref_steps,
                "warmup": warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32 (fp64 gathers)", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["threads"], "kind": "reference",
                                 "sample": info["sample"], "same_config": True, "sample_particles": info["particles"],
                                 "scene_particles": info["scene_particles"],
                                 "stage_s": {k: info[k] for k in ("t_p2g", "t_g2p", "t_advect")}},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from blender_flip_fluids_b200 import engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(vals):
        if world == 1:
            return [int(v) for v in vals]
        t = torch.tensor([int(v) & 0x7FFFFFFFFFFFFFFF for v in vals], device="cuda", dtype=torch.int64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [int(sum(int(p[i]) for p in parts) & 0x7FFFFFFFFFFFFFFF) for i in range(len(vals))]

    # a dedicated (non-default) stream, made torch's current stream BEFORE anything binds to it: the library
    # launches on it (GpuBackend / ffb200_set_stream), NCCL orders its transfers against it, the events time it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    peak, peak_src = measured_peaks()

    run = Runner(n, rank, world, local_rank, stream)
    n_local = run.n_local
    for _ in range(warmup):
        run.step()
    barrier()
    if run.sim is not None and getattr(run.sim, "_profile", False):
        run.sim._phase = {}                   # FFB200_SLAB_PROFILE=1: steady-state phases only
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step = run.time_steps(steps, barrier)
    sampler.stop_flag = True
    sampler.join()
    checks = None
    if not args.no_checksum:
        cnt, ph, fh = run.checksums()
        tot = allsum([cnt, ph, fh or 0])
        checks = {"after_steps": warmup + steps, "particles": tot[0], "particle_hash": f"{tot[1]:016x}", "p2g_field_hash": f"{tot[2]:016x}",
                  "what": "sums mod 2^63 over the owned particles of hash(global id, position bits, velocity bits), and over the "
                          "owned faces of hash(face index) * value bits of the P2G of that state; decomposition independent"}
    stage, launches = run.stage_times()
    if world > 1:
        tt = torch.tensor([ms_per_step, float(n_local)], device="cuda", dtype=torch.float64)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_per_step, n_total = float(mx[0]), int(round(float(sm[1])))
    else:
        n_total = n_local
    value = n_total / (ms_per_step * 1e-3)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "step_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_step")
    roofline = roofline_block(n_total / world, ms_per_step, stage, peak, peak_src, traffic)

    # ---- tolerance mode beside it: fp32 gathers within the north star's 1e-5 (ffb200_set_precision), same evolving batch ----
    tol_rec = None
    if not args.no_tolerance:
        run.ctx.set_precision(True)
        for _ in range(2):
            run.step()
        run.ctx.tolerance_stats(reset=True)
        ms_tol = run.time_steps(steps, barrier)
        ts = run.ctx.tolerance_stats()
        stage_tol, _ = run.stage_times()
        run.ctx.set_precision(False)
        if world > 1:
            tt = torch.tensor([ms_tol, float(ts["advected"]), float(ts["advected_exact"])], device="cuda", dtype=torch.float64)
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms_tol, ts = float(mx[0]), {"advected": int(sm[1]), "advected_exact": int(sm[2])}
        tol_rec = {"ms_per_step": ms_tol, "value": n_total / (ms_tol * 1e-3),
                   "exact_fallback_rate": ts["advected_exact"] / max(1, ts["advected"]),
                   "roofline": roofline_block(n_total / world, ms_tol, stage_tol, peak, peak_src, None),
                   "what": "G2P and RK3 with the trilinear interpolant in fp32 from float-pair cell coordinates (no fp64, no "
                           "conversion instructions); particles whose collision decisions are not clear-cut run the exact code "
                           "(the fallback rate); binning, sort and valid masks are mode independent. Parity: "
                           "tests/test_gpu_parity.py::test_tolerance_mode_* (1e-5 on every fixture and oracle scene). "
                           "`value` above is the EXACT mode."}

    # ---- fixed-batch figure beside it (the round-1 mode: every step re-sorts the same resident batch) ------------
    fixed_rec = None
    if world == 1:
        run.set_fixed(True)
        for _ in range(3):
            run.step()
        ms_fixed = run.time_steps(max(4, steps // 2), barrier)
        fixed_rec = {"ms_per_step": ms_fixed, "value": n_total / (ms_fixed * 1e-3),
                     "batch": "fixed resident batch (each step re-bins, re-sorts and transfers the same particles; results go "
                              "to the spare SoA buffer; the already sorted input is the reorder's best case)"}
        run.set_fixed(False)

    # ---- e2e: host buffers through the reference-facing entry points, particles resident ---------------------------
    e2e = None
    if not args.no_e2e:
        b = run.ctx.device_buffers()
        if world == 1:
            shapes = engine.mac_shapes(n, n, n)
            h_out = [torch.empty(s, dtype=torch.float32).pin_memory() for s in shapes]
            h_valid = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in shapes]
            h_in = [t.cpu().pin_memory() for t in run.tg]
            out = tuple(t.numpy() for t in h_out) + tuple(t.numpy() for t in h_valid)
            mac_in = tuple(t.numpy() for t in h_in)
            ctx = run.ctx

            def e2e_step():
                # the interposed call sites of FluidSimulation::_stepFluid with the particles resident: faces + masks out
                # (the host pressure solve's input), projected field in, advected state stays, CFL speed out
                ctx.declare_resident(particles=True)
                ctx.velocity_advector_advect(None, None, radius=run.radius, method=run.m, out=out)
                ctx.declare_resident(particles=True)
                ctx.update_marker_particle_velocities(None, None, mac_in, method=run.m, ratio_pic_flip=RATIO)
                ctx.declare_resident(particles=True, field=True)
                ctx.advance_marker_particles(None, None, None, None, dt=run.dt, cfl=5.0)
                return ctx.maximum_particle_speed()

            ngrid = sum(int(np.prod(s)) for s in shapes)
            h2d, d2h = ngrid * 4, ngrid * 5 + 8
            path = ("ffb200_velocity_advector_advect (resident particles; faces + valid masks to pinned host) + "
                    "ffb200_update_marker_particle_velocities (projected field from pinned host; results stay resident) + "
                    "ffb200_advance_marker_particles (resident) + ffb200_get_maximum_particle_speed (CFL input to the host)")
        else:
            own = run.sim._owned_field_views()
            h_out = [torch.empty(v.shape, dtype=torch.float32).pin_memory() for v in own]
            h_in = [t.cpu().pin_memory() for t in run.tg]

            def e2e_step():
                run.sim.step_fast(run.radius, RATIO, run.dt, apply_migration=True, projected_field=h_in, p2g_download=h_out,
                                  overlap=run.overlap)
                return run.ctx.maximum_particle_speed()

            h2d = sum(t.numel() for t in h_in) * 4
            d2h = sum(t.numel() for t in h_out) * 4 + 8
            path = ("per rank: SlabSimulation.step_fast with the owned planes of the transferred field copied to pinned host, "
                    "the projected field (stored planes incl. halo) copied from pinned host, particles resident and migrating "
                    "over NCCL, CFL speed read back")
        for _ in range(2):
            e2e_step()
        barrier()
        k_e2e = max(3, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        t_e2e = (time.perf_counter() - t0) / k_e2e
        if world > 1:
            tt = torch.tensor([t_e2e, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            t_e2e, h2d, d2h = float(mx[0]), float(sm[1]), float(sm[2])
        e2e = {"value": n_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": t_e2e * 1e3, "steps": k_e2e, "path": path,
               "particles": "resident across stages and substeps (uploaded once, outside the timed region); what the host "
                            "pipeline needs every substep crosses PCIe inside it"}
    if run.sim is not None and getattr(run.sim, "_profile", False):
        # FFB200_SLAB_PROFILE=1: synchronised wall-clock phases of step_fast (diagnostic; the timed numbers above then
        # include the synchronisation and are not benchmark figures)
        ph = run.sim._phase
        nst = steps + 3
        sys.stderr.write(f"rank {rank} slab phases, ms per step (synchronised): " + ", ".join(f"{k} {1e3 * v / nst:.3f}" for k, v in ph.items()) +
                         f"; total {1e3 * sum(ph.values()) / nst:.3f}; exchange repeats {getattr(run.sim, 'overflows', 0)}\n")
    run_overlap = run.overlap
    class run_sim_overflows:                  # exchanges repeated because a section overflowed (slab.step_fast), this rank
        v = getattr(run.sim, "overflows", 0) if run.sim is not None else None
    run.close()

    # ---- secondary record: BASELINE configs[1], 128^3 APIC (N = 1 only) ----------------------------------------------
    secondary = None
    if world == 1 and not args.no_secondary and n != GRID_SECONDARY:
        try:
            r2 = Runner(GRID_SECONDARY, 0, 1, local_rank, stream)
            rec = {"workload": workload_string(GRID_SECONDARY, r2.n_local)}
            k2 = max(20, steps)
            for mode in ("evolving", "evolving_tolerance", "fixed"):
                r2.set_fixed(mode == "fixed")
                r2.ctx.set_precision(mode == "evolving_tolerance")
                for _ in range(3):
                    r2.step()
                eager = r2.time_steps(k2, barrier)
                st, ln = r2.stage_times()
                graphed = None
                try:
                    g = r2.capture()
                    graphed = r2.time_steps(k2, barrier, replay=g.replay)
                    del g
                except Exception as ex:                                  # capture is an optimisation, never a requirement
                    sys.stderr.write(f"CUDA graph capture unavailable ({ex})\n")
                    torch.cuda.set_stream(stream)
                    torch.cuda.synchronize()
                best = min(eager, graphed) if graphed else eager
                rec[mode] = {"ms_per_step_eager": eager, "ms_per_step_graph": graphed, "value_eager": r2.n_local / (eager * 1e-3),
                             "value_graph": (r2.n_local / (graphed * 1e-3)) if graphed else None, "launches_per_step": ln,
                             "roofline": roofline_block(r2.n_local, best, st, peak, peak_src, None)}
            r2.close()
            secondary = rec
        except Exception as ex:
            secondary = {"error": str(ex)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, info = reference_arm(n, 2, 1)
            cpu = {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "reference", "sample": info["sample"],
                   "same_config": True, "sample_particles": info["particles"], "scene_particles": info["scene_particles"],
                   "cpu_model": host_cpu()[0], "stage_s": {k: info[k] for k in ("t_p2g", "t_g2p", "t_advect")}}
        except Exception as e:                                   # the checker is optional for the headline
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        cfg = dict(base_config,
                   parallelism="single GPU" if world == 1 else f"z-slab x{world} ({n // world} planes per rank, halo {HALO}, ghost "
                               f"particle layers {GHOST}), face halo + particle migration over NCCL, one process per GPU" +
                               (", halo and migrant/ghost exchanges overlapped with the interior particles' G2P + advection" if run_overlap else ""),
                   l2="inputs larger than L2 (particle streams + grids >> 126 MB per step)",
                   batch="evolving: each step sorts and transfers the particles the previous step advected (ranks migrate them); "
                         "the MAC field is replaced by an analytic divergence-free field where the CPU projection would return one",
                   launch="eager launches (every N)",
                   exchange_repeats=(getattr(run_sim_overflows, "v", None)))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (fp64 index/gather arithmetic)", "data": "synthetic", "config": cfg,
                "clocks": sampler.result(), "e2e": e2e, "gpu_launches": launches * steps, "roofline": roofline,
                "checksum": checks, "tolerance_mode": tol_rec, "fixed_batch": fixed_rec, "secondary": secondary,
                "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
