"""Synthetic FLIP scenes: dam break and fill-box particle sets.

Pure numpy, no GPU and no oracle dependency: the same arrays are fed to the CUDA path, to
the C oracle and to the reference harness, so every comparison runs on identical inputs.

Layouts follow the reference's host formats: particle attributes are float32 [N, 3]
(``std::vector<vmath::vec3>``, 12-byte AoS, particlesystem.h:303-325); grids are C-order
[K, J, I(+1)] views of the x-fastest ``Array3d`` storage (array3d.h:774-777).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scene:
    isize: int
    jsize: int
    ksize: int
    dx: float
    pos: np.ndarray                      # float32 [N,3]
    vel: np.ndarray                      # float32 [N,3]
    affx: np.ndarray | None = None       # float32 [N,3] (APIC only)
    affy: np.ndarray | None = None
    affz: np.ndarray | None = None
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    @property
    def radius(self) -> float:
        """P2G particle radius 0.5*sqrt(3)*dx (fluidsimulation.cpp:4351, 5635)."""
        return 0.5 * self.dx * math.sqrt(3.0)


def _seed_cells(i0, i1, j0, j1, k0, k1, dx, ppc, rng, jitter=0.05):
    """Particles for the cell box [i0,i1)x[j0,j1)x[k0,k1), in (k, j, i, site) order.

    ppc=8: the reference's 8 sub-cell sites centre +- dx/4 (fluidsimulation.cpp:8045-8055),
    each jittered by U(-jitter*dx, jitter*dx). ppc=m^3: an m^3 lattice of sites per cell.
    ppc=4: a random 4-subset of the 8 sites (config #5 sweep).
    """
    m = round(ppc ** (1.0 / 3.0))
    sub4 = ppc == 4
    if sub4:
        m = 2
    elif m ** 3 != ppc:
        raise ValueError("ppc must be 4 or a cube (8, 27, ...)")
    kk, jj, ii = np.meshgrid(np.arange(k0, k1), np.arange(j0, j1), np.arange(i0, i1), indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)      # [C,3]
    s = (np.arange(m) + 0.5) / m
    sz, sy, sx = np.meshgrid(s, s, s, indexing="ij")
    sites = np.stack([sx.ravel(), sy.ravel(), sz.ravel()], axis=1)                          # [m^3,3]
    ncell = cells.shape[0]
    if sub4:
        pick = np.argsort(rng.random((ncell, 8)), axis=1)[:, :4]
        pick.sort(axis=1)
        site_xyz = sites[pick]                                                              # [C,4,3]
    else:
        site_xyz = np.broadcast_to(sites[None], (ncell, sites.shape[0], 3))
    p = (cells[:, None, :] + site_xyz) * dx
    p = p + rng.uniform(-jitter * dx, jitter * dx, size=p.shape)
    return p.reshape(-1, 3).astype(np.float32)


def _velocities(n, mode, v0, rng, pos=None, dx=1.0):
    if mode == "random":
        return (rng.uniform(-1.0, 1.0, size=(n, 3)) * v0).astype(np.float32)
    if mode == "swirl":
        # smooth divergence-light field: rotation about the domain's z axis plus a downdraft
        c = pos.astype(np.float64)
        cx, cy = c[:, 0].mean(), c[:, 1].mean()
        v = np.stack([-(c[:, 1] - cy), (c[:, 0] - cx), 0.25 * np.sin(6.0 * c[:, 0])], axis=1) * v0
        v += rng.uniform(-0.05, 0.05, size=v.shape) * v0
        return v.astype(np.float32)
    if mode == "zero":
        return np.zeros((n, 3), np.float32)
    raise ValueError(mode)


def _affine(n, dx, rng, scale=0.1):
    """APIC affine rows ~ U(-1,1)*scale/dx (a velocity gradient of order 0.1 v0 per cell)."""
    return [(rng.uniform(-1.0, 1.0, size=(n, 3)) * (scale / dx)).astype(np.float32) for _ in range(3)]


def dam_break(n: int, ppc: int = 8, apic: bool = False, seed: int = 1234, dx: float | None = None,
              vel: str = "random", v0: float = 1.0, dims: tuple[int, int, int] | None = None) -> Scene:
    """Dam break of SURVEY 8d: fluid cells 3<=i<0.4n, 3<=j<0.8n, 3<=k<n-3 (full z)."""
    I, J, K = dims if dims else (n, n, n)
    dx = (1.0 / n) if dx is None else dx
    rng = np.random.default_rng(seed)
    pos = _seed_cells(3, max(4, int(0.4 * I)), 3, max(4, int(0.8 * J)), 3, K - 3, dx, ppc, rng)
    v = _velocities(pos.shape[0], vel, v0, rng, pos, dx)
    sc = Scene(I, J, K, dx, pos, v, name=f"dam_break_{I}x{J}x{K}_ppc{ppc}_{'apic' if apic else 'flip'}")
    if apic:
        sc.affx, sc.affy, sc.affz = _affine(pos.shape[0], dx, rng)
    sc.meta = dict(kind="dam_break", ppc=ppc, seed=seed, vel=vel, v0=v0)
    return sc


def fill_box(n: int, ppc: int = 8, apic: bool = False, seed: int = 1234, dx: float | None = None,
             vel: str = "random", v0: float = 1.0, fill: float = 0.6) -> Scene:
    """Fill-box of SURVEY 8d (config #3): cells 3<=i,k<n-3, 3<=j<fill*n."""
    dx = (1.0 / n) if dx is None else dx
    rng = np.random.default_rng(seed)
    pos = _seed_cells(3, n - 3, 3, max(4, int(fill * n)), 3, n - 3, dx, ppc, rng)
    v = _velocities(pos.shape[0], vel, v0, rng, pos, dx)
    sc = Scene(n, n, n, dx, pos, v, name=f"fill_box_{n}_ppc{ppc}_{'apic' if apic else 'flip'}")
    if apic:
        sc.affx, sc.affy, sc.affz = _affine(pos.shape[0], dx, rng)
    sc.meta = dict(kind="fill_box", ppc=ppc, seed=seed, vel=vel, v0=v0, fill=fill)
    return sc


def analytic_solid_sdf(I: int, J: int, K: int, dx: float, sphere: tuple[float, float, float, float] | None = None):
    """Node-centred solid SDF phi[(K+1),(J+1),(I+1)] and the 3dx near-solid mask.

    Stand-in for the reference's mesh level set: the domain wall as a solid,
    phi = distance to the boundary box inset (3dx + 1e-4)/2 per side as _getBoundaryAABB
    does (fluidsimulation.cpp:5175-5180; negative outside the box = inside the wall), min'ed
    with an optional sphere obstacle. The near-solid mask follows
    fluidsimulation.cpp:5437-5480: cell (i,j,k) with |phi(node i,j,k)| < 3dx marks coarse cell
    (i//3, j//3, k//3), then ceil(5/3)=2 rounds of 6-neighbour feathering.
    """
    inset = 0.5 * (3.0 * dx + 1e-4)
    z, y, x = np.meshgrid(np.arange(K + 1) * dx, np.arange(J + 1) * dx, np.arange(I + 1) * dx, indexing="ij")
    d = np.minimum.reduce([x - inset, I * dx - inset - x, y - inset, J * dx - inset - y, z - inset, K * dx - inset - z])
    phi = d
    if sphere is not None:
        cx, cy, cz, r = sphere
        phi = np.minimum(phi, np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) - r)
    phi = phi.astype(np.float32)
    gi, gj, gk = (math.ceil(I / 3), math.ceil(J / 3), math.ceil(K / 3))
    near = np.zeros((gk, gj, gi), np.uint8)
    band = np.abs(phi[:K, :J, :I]) < np.float32(3.0 * dx)
    kk, jj, ii = np.nonzero(band)
    near[kk // 3, jj // 3, ii // 3] = 1
    for _ in range(2):
        g = near.copy()
        g[1:, :, :] |= near[:-1, :, :]
        g[:-1, :, :] |= near[1:, :, :]
        g[:, 1:, :] |= near[:, :-1, :]
        g[:, :-1, :] |= near[:, 1:, :]
        g[:, :, 1:] |= near[:, :, :-1]
        g[:, :, :-1] |= near[:, :, 1:]
        near = g
    return phi, near


# ---- decomposition-independent generator (large scenes, any rank count) ------------------------------
#
# The jitter / velocity / affine values of a particle are a pure function of (seed, global particle
# id, component): a counter-based integer hash, evaluated with the same integer and float64
# operations by numpy on the host and by torch on the device, so a z-slab rank, a single GPU and
# the CPU reference arm all produce bit-identical particles for the planes they hold without ever
# materialising the whole scene in one place.

_SITE8 = [(0.25, 0.25, 0.25), (0.75, 0.25, 0.25), (0.25, 0.75, 0.25), (0.75, 0.75, 0.25),
          (0.25, 0.25, 0.75), (0.75, 0.25, 0.75), (0.25, 0.75, 0.75), (0.75, 0.75, 0.75)]
HASH_COMPONENTS = 16          # 3 jitter + 3 velocity + 9 affine, padded


def _hash_u01(xp, ids, comp, seed):
    """lowbias32-style integer hash of (seed, id, comp) -> float64 in [0, 1) with 24 random bits.
    `ids` is an int64 array/tensor (numpy or torch); all intermediates are int64 holding 32-bit values."""
    M = 0xFFFFFFFF
    lo = ids & M
    hi = (ids >> 32) & M
    x = (lo * 0x9E3779B1 + hi * 0x7F4A7C15 + (comp * 0x85EBCA77 + seed * 0xC2B2AE3D + 0x27D4EB2F)) & M
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & M
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & M
    x = x ^ (x >> 16)
    return (x >> 8).to(xp.float64) * (1.0 / 16777216.0) if xp.__name__ == "torch" else (x >> 8).astype(np.float64) * (1.0 / 16777216.0)


def dam_break_extent(I, J, K):
    """Fluid cell box of the SURVEY 8d dam break: [3, 0.4 I) x [3, 0.8 J) x [3, K - 3)."""
    return 3, max(4, int(0.4 * I)), 3, max(4, int(0.8 * J)), 3, K - 3


def dam_break_count(I, J, K, ppc=8, k0=None, k1=None):
    i0, i1, j0, j1, kk0, kk1 = dam_break_extent(I, J, K)
    k0 = kk0 if k0 is None else max(kk0, k0)
    k1 = kk1 if k1 is None else min(kk1, k1)
    return max(0, k1 - k0) * (j1 - j0) * (i1 - i0) * ppc


def dam_break_planes(I, J, K, dx, k0, k1, apic=True, v0=0.5, seed=1234, jitter=0.05, xp=np, device=None, chunk_planes=8):
    """Particles of the cell planes [k0, k1) of the hashed dam break I x J x K (8 per cell, the reference's
    sub-cell sites, fluidsimulation.cpp:8045-8055), as SoA float32 streams [px,py,pz,vx,vy,vz(,9 affine)]
    plus int64 global ids, in (k, j, i, site) order. xp = numpy (host) or torch (`device`)."""
    i0, i1, j0, j1, kk0, kk1 = dam_break_extent(I, J, K)
    k0, k1 = max(kk0, k0), min(kk1, k1)
    ni, nj = i1 - i0, j1 - j0
    per_plane = ni * nj * 8
    nstream = 15 if apic else 6
    is_torch = xp.__name__ == "torch"
    n = max(0, k1 - k0) * per_plane
    if is_torch:
        out = [xp.empty(n, dtype=xp.float32, device=device) for _ in range(nstream)]
        ids_out = xp.empty(n, dtype=xp.int64, device=device)
        ar = lambda m: xp.arange(m, dtype=xp.int64, device=device)
        site = xp.tensor(_SITE8, dtype=xp.float64, device=device)
        f64 = lambda t: t.to(xp.float64)
        f32 = lambda t: t.to(xp.float32)
    else:
        out = [np.empty(n, np.float32) for _ in range(nstream)]
        ids_out = np.empty(n, np.int64)
        ar = lambda m: np.arange(m, dtype=np.int64)
        site = np.array(_SITE8, np.float64)
        f64 = lambda t: t.astype(np.float64)
        f32 = lambda t: t.astype(np.float32)
    for ka in range(k0, k1, chunk_planes):
        kb = min(k1, ka + chunk_planes)
        m = (kb - ka) * per_plane
        local = ar(m)
        s = local % 8
        c = local // 8
        ci = c % ni + i0
        cj = (c // ni) % nj + j0
        ck = c // (ni * nj) + ka
        gid = local + (ka - kk0) * per_plane
        o = (ka - k0) * per_plane
        cell = (ci, cj, ck)
        for a in range(3):
            u = _hash_u01(xp, gid, a, seed)
            p = (f64(cell[a]) + site[s, a]) * dx + (2.0 * u - 1.0) * (jitter * dx)
            out[a][o:o + m] = f32(p)
        for a in range(3):
            u = _hash_u01(xp, gid, 3 + a, seed)
            out[3 + a][o:o + m] = f32((2.0 * u - 1.0) * v0)
        if apic:
            for a in range(9):
                u = _hash_u01(xp, gid, 6 + a, seed)
                out[6 + a][o:o + m] = f32((2.0 * u - 1.0) * (0.1 / dx))
        ids_out[o:o + m] = gid
    return out, ids_out


def analytic_solid_sdf_planes(I, J, K, dx, k0, k1, xp=np, device=None):
    """Node planes [k0, k1] (inclusive) of analytic_solid_sdf's wall SDF (no sphere), float32 [(k1-k0+1), J+1, I+1]:
    the same double expression, narrowed once, so slabs and the full array agree bit for bit."""
    inset = 0.5 * (3.0 * dx + 1e-4)
    if xp.__name__ == "torch":
        ar = lambda a, b: xp.arange(a, b, dtype=xp.float64, device=device)
        z, y, x = xp.meshgrid(ar(k0, k1 + 1) * dx, ar(0, J + 1) * dx, ar(0, I + 1) * dx, indexing="ij")
        d = xp.minimum(xp.minimum(xp.minimum(x - inset, I * dx - inset - x), xp.minimum(y - inset, J * dx - inset - y)),
                       xp.minimum(z - inset, K * dx - inset - z))
        return d.to(xp.float32)
    z, y, x = np.meshgrid(np.arange(k0, k1 + 1) * dx, np.arange(J + 1) * dx, np.arange(I + 1) * dx, indexing="ij")
    d = np.minimum.reduce([x - inset, I * dx - inset - x, y - inset, J * dx - inset - y, z - inset, K * dx - inset - z])
    return d.astype(np.float32)
