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
