#!/usr/bin/env python
"""End-to-end through the DROP-IN: the same simulation driven through the reference's own FluidSimulation_* C ABI,
once with the unmodified reference library (all host threads) and once with libffengine_b200.so (the reference's
pageable std::vector / Array3d containers, page-locked by the interposer; particles resident; everything outside the
interposed stages -- pressure solve, level sets, ... -- still reference CPU code in both runs).

    python tools/bench_dropin.py [--n 128] [--frames 3] [--method apic]

Per frame: wall time, substeps and the reference's own timing block (FluidSimulation_get_frame_stats_data:
`advection` = the "Advect Velocity Field" stage = P2G + extrapolation, `particles` = G2P + advection + removal), so the
interposed stages are timed by the reference's own timers in both runs. The first frame of the drop-in holds the CUDA
context creation and every first-use allocation and is listed separately. One JSON line on stdout."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libffengine_ref.so")
DROPIN = os.path.join(ROOT, "blender_flip_fluids_b200", "lib", "libffengine_b200.so")


def child(lib, n, frames, method):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from blender_flip_fluids_b200 import scenes
    from ffengine_mini import Engine
    sc = scenes.dam_break(n, apic=(method == "apic"), vel="swirl", v0=0.4, seed=17)
    e = Engine(lib, n, n, n, sc.dx)
    e.disable_console_output()
    e.disable_surface_reconstruction()
    if method == "apic":
        e.set_apic()
    e.set_picflip_ratio(0.05)
    e.set_max_thread_count(os.cpu_count())
    e.add_body_force(0.0, -9.81, 0.0)
    e.load_marker_particle_data(sc.pos, sc.vel)
    if method == "apic":
        e.load_marker_particle_affine_data(sc.affx * 0.01, sc.affy * 0.01, sc.affz * 0.01)
    e.initialize()
    out = []
    for _ in range(frames):
        t0 = time.perf_counter()
        e.update(1.0 / 60.0)
        wall = time.perf_counter() - t0
        st = e.frame_stats()
        out.append({"wall_s": wall, "substeps": st.substeps, "fluid_particles": st.fluid_particles,
                    "timing": {k: getattr(st.timing, k) for k, _ in st.timing._fields_}})
    e.close()
    print("FRAMES " + json.dumps(out))


def run(lib, n, frames, method, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", lib, "--n", str(n), "--frames", str(frames), "--method", method],
                       capture_output=True, text=True, env=env, timeout=3000)
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("FRAMES ")][-1][7:])


def summarize(frames, skip):
    use = frames[skip:]
    sub = sum(f["substeps"] for f in use)
    upd = sum(f["substeps"] * f["fluid_particles"] for f in use)
    stage = sum(f["timing"]["advection"] + f["timing"]["particles"] for f in use)
    return {"frames": len(use), "substeps": sub, "particle_updates": upd, "wall_s": sum(f["wall_s"] for f in use),
            "interposed_stage_s": stage, "interposed_stage_ms_per_substep": 1e3 * stage / max(sub, 1),
            "particle_updates_per_s_in_interposed_stages": upd / stage if stage > 0 else None,
            "timing_s": {k: sum(f["timing"][k] for f in use) for k in use[0]["timing"]} if use else {}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child")
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--method", default="apic")
    a = ap.parse_args()
    if a.child:
        return child(a.child, a.n, a.frames, a.method)
    prof = os.path.join("/tmp", f"ffb200_dropin_profile_{os.getpid()}.json")
    ref = run(REF, a.n, a.frames, a.method)
    got = run(DROPIN, a.n, a.frames, a.method, {"FFB200_DROPIN_PROFILE": prof})
    stages = None
    if os.path.exists(prof):
        with open(prof) as f:
            stages = json.load(f)
        os.remove(prof)
    res = {"what": f"dam break {a.n}^3, {a.method}, {ref[0]['fluid_particles']} particles, {a.frames} frames of 1/60 s through FluidSimulation_update; "
                   "frame 1 listed apart (CUDA context + first-use allocations in the drop-in run)",
           "host_threads": os.cpu_count(),
           "reference": {"first_frame": summarize(ref, 0) if a.frames == 1 else summarize(ref[:1], 0), "steady": summarize(ref, 1)},
           "dropin": {"first_frame": summarize(got[:1], 0), "steady": summarize(got, 1), "interposer_stage_profile_all_frames": stages}}
    r, g = res["reference"]["steady"], res["dropin"]["steady"]
    if r["interposed_stage_s"] and g["interposed_stage_s"]:
        res["steady_speedup_interposed_stages"] = r["interposed_stage_ms_per_substep"] / g["interposed_stage_ms_per_substep"]
        res["steady_speedup_whole_frame"] = (r["wall_s"] / max(r["substeps"], 1)) / (g["wall_s"] / max(g["substeps"], 1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
