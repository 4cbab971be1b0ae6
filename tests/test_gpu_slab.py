"""Two-rank z-slab parity (-m gpu): the slab driver with the CUDA plumbing kernels on two ranks
against ONE single-GPU context on the same particles -- ids, positions, velocities and affine rows
bit-identical after two substeps with migration (also with undersized exchange sections, which are repeated, and with
the exchange overlapped with the interior particles' G2P + advection). With two or more devices the ranks use one GPU each
over NCCL; on a single-GPU box both ranks share device 0 and exchange over gloo (staged through
host memory by slab._sendrecv), which still runs every slab kernel."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from blender_flip_fluids_b200 import slab, scenes, engine
out, apic = sys.argv[2], sys.argv[3] == "apic"
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
shared = torch.cuda.device_count() < world
if shared:
    lr = 0
torch.cuda.set_device(lr)
if shared:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
I, J, K, dx = 32, 24, 64, 0.01
sc = scenes.dam_break(32, apic=apic, dx=dx, dims=(I, J, K), vel="random", v0=0.4, seed=33)
sc.vel[:, 2] += np.where(sc.pos[:, 2] < 0.5 * K * dx, 0.5, -0.5).astype(np.float32)
phi, near = scenes.analytic_solid_sdf(I, J, K, dx)
kb, ke = slab.slab_range(K, world, rank)
be = slab.GpuBackend(I, J, K, dx, kb, ke, 7, lr, apic)
be.set_solid(phi, near)
sim = slab.SlabSimulation(I, J, K, dx, rank, world, be, halo=7, ghost=1 if sys.argv[4] == "fast-tiny" else 2)
if sys.argv[4] == "fast-tiny":
    sim.force_initial_caps = (8, 16, 8, 16)       # every merged exchange overflows and is repeated with resized sections
kz = np.floor(sc.pos[:, 2].astype(np.float64) * (1.0 / dx)).astype(np.int64)
sel = np.nonzero((kz >= kb) & (kz < ke))[0]
cols = [sc.pos[sel, 0], sc.pos[sel, 1], sc.pos[sel, 2], sc.vel[sel, 0], sc.vel[sel, 1], sc.vel[sel, 2]]
if apic:
    for a in (sc.affx, sc.affy, sc.affz):
        cols += [a[sel, 0], a[sel, 1], a[sel, 2]]
dev = torch.device("cuda", lr)
sim.set_particles([torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in cols], torch.from_numpy(sel.astype(np.int32)).to(dev))
dt = 1.5 * dx / 0.9
if sys.argv[4].startswith("fast"):
    sim.load_resident()
    for _ in range(2):
        sim.step_fast(sc.radius, 0.05, dt, overlap=(sys.argv[4] == "fast-overlap"))
    sim.sync_from_backend()
    assert sys.argv[4] != "fast-tiny" or getattr(sim, "overflows", 0) >= 2
else:
    for _ in range(2):
        sim.step(sc.radius, 0.05, dt)
allp, ids = sim.gather_particles()
if rank == 0:
    # single-GPU run of the same two substeps on this rank's device
    m = engine.APIC if apic else engine.FLIP
    with engine.FlipContext(I, J, K, dx, device=lr) as ctx:
        ctx.set_solid(phi, near)
        ctx.set_particles(sc.pos, sc.vel, sc.affx, sc.affy, sc.affz)
        for _ in range(2):
            ctx.p2g(sc.radius, m); ctx.save_velocity_field(); ctx.g2p(m, 0.05); ctx.advect(dt, 5.0, True)
        p, v, ax, ay, az = ctx.get_particles(pos=True, vel=True, affine=apic)
    want = [p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2]]
    if apic:
        for a in (ax, ay, az):
            want += [a[:, 0], a[:, 1], a[:, 2]]
    np.savez(out, got=allp.cpu().numpy(), ids=ids.cpu().numpy(), want=np.stack(want, 0), n=sc.n)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("plumbing", ["fast", "generic", "fast-tiny", "fast-overlap"])
@pytest.mark.parametrize("method", ["flip", "apic"])
def test_two_gpu_slab_matches_single_gpu(tmp_path, method, plumbing):
    import torch
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "res.npz")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT, out, method, plumbing],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    z = np.load(out)
    assert np.array_equal(z["ids"], np.arange(int(z["n"])))
    assert z["got"].tobytes() == z["want"].astype(np.float32).tobytes()
