"""Oracle-backed CPU stand-in for blender_flip_fluids_b200.slab.GpuBackend -- TEST INFRASTRUCTURE.

Lets the slab driver's exchange logic (ghost particles, face halos, migration) run under gloo
on CPU tensors. Grids are kept full-size but only the rank's STORED planes are ever non-zero,
and after P2G everything outside the OWNED planes is zeroed, so a missing or misplaced halo
exchange changes the result.
"""
import numpy as np
import torch

from oracle import flip_oracle as fo


class CpuOracleBackend:
    def __init__(self, I, J, K, dx, k_begin, k_end, halo, apic):
        self.I, self.J, self.K, self.dx = I, J, K, dx
        self.kb, self.ke, self.halo = k_begin, k_end, halo
        self.kbase = max(0, k_begin - halo)
        self.ktop = min(K, k_end + halo)
        self.apic = apic
        self.device = torch.device("cpu")
        self.streams, self.ids = None, None
        shp = fo.mac_shapes(I, J, K)
        self.field = [torch.zeros(s, dtype=torch.float32) for s in shp]
        self.saved = [torch.zeros(s, dtype=torch.float32) for s in shp]
        self.phi, self.near = None, None

    def load_particles(self, streams, ids):
        self.streams = [s.clone() for s in streams]
        self.ids = ids.clone()

    def particle_views(self, n=None):
        return self.streams, self.ids

    def _np(self, lo, hi):
        return np.stack([s.numpy() for s in self.streams[lo:hi]], axis=1).astype(np.float32)

    def p2g(self, radius):
        order = np.argsort(self.ids.numpy().astype(np.int64) & 0xffffffff, kind="stable")   # reference sums by index
        pos, vel = self._np(0, 3)[order], self._np(3, 6)[order]
        aff = [self._np(6 + 3 * d, 9 + 3 * d)[order] for d in range(3)] if self.apic else [None] * 3
        (u, v, w), _ = fo.p2g(self.I, self.J, self.K, self.dx, radius, fo.APIC if self.apic else fo.FLIP, pos, vel, *aff)
        for d, a in enumerate((u, v, w)):
            t = torch.from_numpy(a)
            own_hi = self.ke + (1 if (d == 2 and self.ke == self.K) else 0)     # top rank also owns w plane K
            t[:self.kb] = 0
            t[own_hi:] = 0
            self.field[d].copy_(t)

    def field_planes(self, d, saved=False):
        f = (self.saved if saved else self.field)[d]
        kstore = (self.ktop - self.kbase) + (1 if d == 2 else 0)
        return f[self.kbase:self.kbase + kstore].reshape(kstore, -1), self.kbase

    def save_field(self):
        for d in range(3):
            self.saved[d].copy_(self.field[d])

    def set_solid(self, phi, near):
        self.phi, self.near = phi, near

    def g2p(self, ratio):
        mac = [f.numpy() for f in self.field]
        pos = self._np(0, 3)
        if self.apic:
            vel, ax, ay, az = fo.g2p_apic(self.I, self.J, self.K, self.dx, pos, mac)
            cols = np.concatenate([vel, ax, ay, az], axis=1)
        else:
            cols = fo.g2p_flip(self.I, self.J, self.K, self.dx, pos, self._np(3, 6), mac, [f.numpy() for f in self.saved], ratio)
        for q in range(cols.shape[1]):
            self.streams[3 + q] = torch.from_numpy(np.ascontiguousarray(cols[:, q]))

    def advect(self, dt, cfl=5.0, collide=True):
        out = fo.advect(self.I, self.J, self.K, self.dx, self._np(0, 3), [f.numpy() for f in self.field], self.phi,
                        self.near, dt, cfl, collide)
        for q in range(3):
            self.streams[q] = torch.from_numpy(np.ascontiguousarray(out[:, q]))
