#!/usr/bin/env python
"""Quick single-GPU probe: hashed dam break n^3 through the resident stages, stage times + memory.
    python tools/try_scene.py 512 [steps] [fixed|evolving]"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blender_flip_fluids_b200 import engine, benchscene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mode = sys.argv[3] if len(sys.argv) > 3 else "evolving"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
dx, v0 = 1.0 / n, 0.5
t0 = time.time()
ctx = engine.FlipContext(n, n, n, dx, device=0)
ctx.set_stream(stream.cuda_stream)
benchscene.set_wall_solid(ctx, n, n, n, dx, dev)
N = benchscene.fill_dam_break(ctx, n, n, n, dx, 0, n, True, v0, 1234, dev)
torch.cuda.synchronize()
print("setup s", time.time() - t0, "particles", N, "mem GB", torch.cuda.mem_get_info()[0] / 1e9, "free of", torch.cuda.mem_get_info()[1] / 1e9, flush=True)
radius, dt = 0.5 * dx * math.sqrt(3.0), dx / v0
tg = benchscene.taylor_green_field(n, n, n, dx, 0, n, v0, dev)
fv = benchscene.field_views(ctx, dev)
ctx.set_fixed_batch(mode == "fixed")

def step():
    ctx.p2g(radius, engine.APIC)
    ctx.save_velocity_field()
    if mode != "fixed":
        for a, b in zip(fv, tg):
            a.copy_(b.view(-1))
    ctx.g2p(engine.APIC, 0.05)
    ctx.advect(dt, 5.0, True)

for _ in range(3):
    step()
torch.cuda.synchronize()
print("after warmup mem free GB", torch.cuda.mem_get_info()[0] / 1e9, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    step()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
step(); t = ctx.timing()
print(json.dumps({"n": n, "mode": mode, "particles": N, "ms_per_step": ms, "Gpps": N / ms / 1e6, "stages": {k: v for k, v in t.items() if k.endswith("_ms")},
                  "checksum": benchscene.particle_checksum(ctx, True, dev), "maxspeed": ctx.maximum_particle_speed()}))
