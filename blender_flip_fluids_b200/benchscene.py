"""Device-side loading of the large synthetic scenes (bench.py, the full-size tests): particles of a
z-range of the hashed dam break generated with torch on the GPU straight into the library's resident
streams, the wall solid SDF and its near-solid grid built on the device. torch is plumbing here
(device memory + elementwise generation of INPUTS); no stage of the hot path runs through it."""
from __future__ import annotations

import math

import torch

from . import scenes
from .slab import _view


def near_solid_from_wall_sdf(I, J, K, dx, device, chunk=24):
    """The 3dx near-solid byte grid of scenes.analytic_solid_sdf (fluidsimulation.cpp:5437-5480) for the
    wall-only SDF, built plane chunk by plane chunk on `device`: cell (i,j,k) with |phi(node i,j,k)| < 3dx marks
    coarse cell (i//3, j//3, k//3); then two rounds of 6-neighbour feathering."""
    gi, gj, gk = math.ceil(I / 3), math.ceil(J / 3), math.ceil(K / 3)
    near = torch.zeros((gk, gj, gi), dtype=torch.uint8, device=device)
    lim = torch.tensor(3.0 * dx, dtype=torch.float32, device=device)      # np.float32(3.0 * dx)
    for k0 in range(0, K, chunk - chunk % 3 if chunk >= 3 else 3):
        k1 = min(K, k0 + (chunk - chunk % 3 if chunk >= 3 else 3))
        phi = scenes.analytic_solid_sdf_planes(I, J, K, dx, k0, k1 - 1, xp=torch, device=device)   # nodes k0..k1-1
        band = (phi[:, :J, :I].abs() < lim)
        kk, jj, ii = torch.nonzero(band, as_tuple=True)
        near[(kk + k0) // 3, jj // 3, ii // 3] = 1
    for _ in range(2):
        g = near.clone()
        g[1:, :, :] |= near[:-1, :, :]
        g[:-1, :, :] |= near[1:, :, :]
        g[:, 1:, :] |= near[:, :-1, :]
        g[:, :-1, :] |= near[:, 1:, :]
        g[:, :, 1:] |= near[:, :, :-1]
        g[:, :, :-1] |= near[:, :, 1:]
        near = g
    return near.contiguous()


def set_wall_solid(ctx, I, J, K, dx, device):
    """Upload the wall-only solid SDF (the context's stored node planes) and the near-solid grid, both built
    on the device (ffb200_set_solid_device)."""
    b = ctx.device_buffers()
    phi = scenes.analytic_solid_sdf_planes(I, J, K, dx, b.kbase, b.kbase + b.kloc, xp=torch, device=device).contiguous()
    near = near_solid_from_wall_sdf(I, J, K, dx, device)
    ctx.set_solid_device(phi.data_ptr(), near.data_ptr())
    ctx.synchronize()
    return phi, near


def fill_dam_break(ctx, I, J, K, dx, k_begin, k_end, apic, v0, seed, device, headroom=1.15, chunk_planes=8):
    """Generate the planes [k_begin, k_end) of the hashed dam break on `device` directly into the context's resident
    SoA streams (ids = global particle ids). Returns the particle count."""
    n = scenes.dam_break_count(I, J, K, 8, k_begin, k_end)
    ctx.reserve_particles(int(n * headroom) + 4096, apic)
    b = ctx.device_buffers()
    ptrs = list(b.pos) + list(b.vel) + (list(b.aff) if apic else [])
    views = [_view(p, n, "<f4", device) for p in ptrs]
    ids = _view(b.ids, n, "<u4", device)
    e = scenes.dam_break_extent(I, J, K)
    k0, k1 = max(e[4], k_begin), min(e[5], k_end)
    per_plane = scenes.dam_break_count(I, J, K, 8, e[4], e[4] + 1)
    for ka in range(k0, k1, chunk_planes):
        kb = min(k1, ka + chunk_planes)
        streams, gid = scenes.dam_break_planes(I, J, K, dx, ka, kb, apic=apic, v0=v0, seed=seed, xp=torch, device=device,
                                               chunk_planes=chunk_planes)
        o = (ka - k0) * per_plane
        m = int(gid.shape[0])
        for v, s in zip(views, streams):
            v[o:o + m] = s
        ids[o:o + m] = gid.to(torch.int32)
        del streams, gid
    ctx.set_num_particles(n, apic)
    return n


def taylor_green_field(I, J, K, dx, kbase, kloc, amp, device):
    """A divergence-free MAC field whose normal component vanishes on the walls of the INNER box [a, 1 - a]^3,
    a = 3 dx (the fluid's margin; the collision surface of the wall solid sits at 1.7 dx), sampled at the face centres
    of the stored planes [kbase, kbase + kloc (+1)): two superposed Taylor-Green vortices in the inner coordinates
    s = clamp((x - a) / (1 - 2a), 0, 1),
        u =  A sin(pi sx) cos(pi sy)
        v = -A cos(pi sx) sin(pi sy) + B sin(pi sy) cos(pi sz)
        w =                          - B cos(pi sy) sin(pi sz),      A = B = amp / 2
    (inside the inner box its discrete divergence on the MAC grid vanishes identically). Stand-in for the
    pressure-projected field in the evolving-batch benchmark: particles circulate in x-y AND across z -- they cross
    slab faces, so ranks really migrate them --, the set does not compress, and nothing is driven into the walls (a
    field tangential only at the DOMAIN boundary packs hundreds of particles per cell against the solid within a few
    steps, which no projected field of the real pipeline does)."""
    f64 = dict(dtype=torch.float64, device=device)
    A = B = 0.5 * amp
    pi = math.pi
    a, L = 3.0 * dx, dx * max(I, J, K) - 6.0 * dx
    inner = lambda t: torch.clamp((t - a) / L, 0.0, 1.0)
    xf = inner(torch.arange(I + 1, **f64) * dx)                  # face / centre coordinates per axis, inner-box units
    xc = inner((torch.arange(I, **f64) + 0.5) * dx)
    yf = inner(torch.arange(J + 1, **f64) * dx)
    yc = inner((torch.arange(J, **f64) + 0.5) * dx)
    zc = inner((torch.arange(kbase, kbase + kloc, **f64) + 0.5) * dx)
    zf = inner(torch.arange(kbase, kbase + kloc + 1, **f64) * dx)
    u2 = (A * torch.sin(pi * xf)[None, :] * torch.cos(pi * yc)[:, None]).to(torch.float32)
    u = u2[None].expand(kloc, J, I + 1).contiguous()
    v = (-A * torch.cos(pi * xc)[None, None, :] * torch.sin(pi * yf)[None, :, None] +
         B * torch.sin(pi * yf)[None, :, None] * torch.cos(pi * zc)[:, None, None]).to(torch.float32).expand(kloc, J + 1, I).contiguous()
    w = (-B * torch.cos(pi * yc)[None, :, None] * torch.sin(pi * zf)[:, None, None]).to(torch.float32).expand(kloc + 1, J, I).contiguous()
    return u, v, w


def field_views(ctx, device, saved=False):
    b = ctx.device_buffers()
    src = b.saved if saved else b.field
    return [_view(src[d], b.face_count[d], "<f4", device) for d in range(3)]


def particle_checksum(ctx, apic, device):
    """Order-independent 64-bit checksum of the owned resident particles: sum over particles of a hash of
    (global id, bit patterns of position and velocity). Equal on any decomposition iff every particle carries
    the same bits. Returns (count, checksum) as python ints (per rank; add across ranks mod 2^63)."""
    b = ctx.device_buffers()
    n = b.n
    if n == 0:
        return 0, 0
    ids = _view(b.ids, n, "<u4", device)
    own = ids >= 0                                       # ghost copies carry the top id bit
    acc = (ids.to(torch.int64) & 0x7FFFFFFF) * 0x9E3779B1
    for q, p in enumerate(list(b.pos) + list(b.vel)):
        bits = _view(p, n, "<f4", device).view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        acc = (acc * 1000003 + bits * (2 * q + 3)) & 0x7FFFFFFFFFFF
    acc = torch.where(own, acc, torch.zeros_like(acc))
    return int(own.sum().item()), int(acc.sum().item() & 0x7FFFFFFFFFFFFFFF)


def field_checksum(ctx, device, k_lo, k_hi, top_w=False):
    """Position-weighted 64-bit checksum of the velocity-field bits on the cell planes [k_lo, k_hi) this rank owns
    (w: face planes [k_lo, k_hi); top_w adds w's last plane K, which belongs to the last rank)."""
    b = ctx.device_buffers()
    total = 0
    for d in range(3):
        plane = b.face_plane[d]
        f = _view(b.field[d], b.face_count[d], "<f4", device).view(-1, plane)
        k_top = k_hi + (1 if (d == 2 and top_w) else 0)
        lo, hi = k_lo - b.kbase, k_top - b.kbase
        bits = f[lo:hi].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        gidx = (torch.arange(k_lo, k_top, dtype=torch.int64, device=device)[:, None] * plane +
                torch.arange(plane, dtype=torch.int64, device=device)[None, :])
        total += int((((gidx * 0x9E3779B1 + (d + 1)) & 0xFFFFFFFF) * bits & 0x7FFFFFFFFFFF).sum().item())
    return total & 0x7FFFFFFFFFFFFFFF
