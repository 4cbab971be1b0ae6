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


class CpuOracleFastBackend(CpuOracleBackend):
    """The device-side plumbing interface of GpuBackend (packed record blocks with device-written
    headers, ghost-bit ids, merged migrant + ghost routing) restated in numpy, so that
    SlabSimulation.step_fast -- its per-face buffer sizing, header handling and exchange protocol --
    runs under gloo on CPU. Mirrors ffb200_slab.cu: k_pack_layers, k_route_mark_ghosts,
    k_fill_collect/k_fill_move (as a plain filter), k_append."""
    fast = True
    GHOST = np.int32(-2 ** 31)

    @property
    def ctx(self):
        return self

    @property
    def n(self):
        return int(self.ids.shape[0])

    def record_floats(self):
        return len(self.streams) + 1

    def reserve(self, capacity):
        pass

    def new_block(self, capacity):
        return torch.zeros(capacity * self.record_floats() + 4, dtype=torch.float32)

    def new_block2(self, cap_m, cap_g):
        return torch.zeros((cap_m + cap_g) * self.record_floats() + 8, dtype=torch.float32)

    def _cells(self):
        return np.floor(self.streams[2].numpy().astype(np.float64) * (1.0 / self.dx)).astype(np.int64)

    def _records(self, sel):
        cols = [s.numpy()[sel] for s in self.streams] + [self.ids.numpy()[sel].view(np.float32)]
        return np.stack(cols, axis=1).astype(np.float32) if len(sel) else np.zeros((0, self.record_floats()), np.float32)

    def _write(self, block, first, cap, sel):
        rows = self.record_floats()
        sel = sel[:cap]                                          # records beyond the capacity are dropped (overflow flag)
        rec = self._records(sel)
        block[first * rows:(first + len(sel)) * rows] = torch.from_numpy(rec.reshape(-1))

    def pack_layers(self, lo_a, hi_a, block_a, lo_b, hi_b, block_b, capacity):
        k = self._cells()
        rows = self.record_floats()
        for lo, hi, blk in ((lo_a, hi_a, block_a), (lo_b, hi_b, block_b)):
            if blk is None:
                continue
            sel = np.nonzero((k >= lo) & (k < hi))[0]
            self._write(blk, 0, capacity, sel)
            blk[capacity * rows:].view(torch.int32)[:] = torch.tensor([len(sel), int(len(sel) > capacity), 0, 0], dtype=torch.int32)

    def _append(self, rec, as_ghost):
        ns = len(self.streams)
        for q in range(ns):
            self.streams[q] = torch.cat([self.streams[q], torch.from_numpy(np.ascontiguousarray(rec[:, q]))])
        ids = np.ascontiguousarray(rec[:, ns]).view(np.int32)
        if as_ghost:
            ids = ids | self.GHOST
        self.ids = torch.cat([self.ids, torch.from_numpy(ids.copy())])

    def append(self, block, count, as_ghost=False):
        self.append_records(block, 0, count, as_ghost)

    def append_records(self, block, first_record, count, as_ghost=False):
        if count:
            rows = self.record_floats()
            rec = block[first_record * rows:(first_record + count) * rows].numpy().reshape(count, rows)
            self._append(rec, as_ghost)

    def route_ghosts_begin(self, k_begin, k_end, g, block_up, block_down, caps):
        rows = self.record_floats()
        k = self._cells()
        ids = self.ids.numpy()
        ghost = ids < 0
        up = ~ghost & (k >= k_end) & (block_up is not None)
        down = ~ghost & (k < k_begin) & (block_down is not None)
        owned_after = ~ghost & ~(up | down)
        gup = owned_after & (block_up is not None) & (k >= k_end - g)
        gdown = owned_after & (block_down is not None) & (k < k_begin + g)
        keep = (up & (k < k_end + g)) | (down & (k >= k_begin - g))
        sent = np.zeros_like(up)
        for m, cap in ((up, caps[0]), (down, caps[2])):
            idx = np.nonzero(m)[0]
            sent[idx[:cap]] = True                               # a migrant that does not fit stays with its sender
        keep &= sent
        leave = ghost | (sent & ~keep)
        for blk, mig, gh, cap_m, cap_g in ((block_up, up, gup, caps[0], caps[1]), (block_down, down, gdown, caps[2], caps[3])):
            if blk is None:
                continue
            self._write(blk, 0, cap_m, np.nonzero(mig)[0])
            self._write(blk, cap_m, cap_g, np.nonzero(gh)[0])
            nm, ng = int(mig.sum()), int(gh.sum())
            blk[(cap_m + cap_g) * rows:].view(torch.int32)[:] = torch.tensor([nm, int(nm > cap_m), ng, int(ng > cap_g), int(leave.sum()), 0, 0, 0],
                                                                               dtype=torch.int32)
        self._keep = keep                                        # ghost bit applied by route_end (the marking may be repeated)
        self._leave = leave
        self._counts = (int((up & sent).sum()), int((down & sent).sum()))

    def route_end(self, leaving=None):
        assert leaving is None or leaving == int(self._leave.sum())
        if getattr(self, "_keep", None) is not None:
            ids = self.ids.numpy().copy()
            ids[self._keep] |= self.GHOST                        # sent, and kept here as the new owner's ghost copy
            self.ids = torch.from_numpy(ids)
            self._keep = None
        stay = torch.from_numpy(np.nonzero(~self._leave)[0])
        self.streams = [s.index_select(0, stay) for s in self.streams]
        self.ids = self.ids.index_select(0, stay)
        return self.n, self._counts[0], self._counts[1]

    def p2g(self, radius):
        saved = self.ids
        self.ids = torch.from_numpy(self.ids.numpy() & np.int32(0x7fffffff))     # sum order = global particle index
        try:
            super().p2g(radius)
        finally:
            self.ids = saved

    def halo_plan(self, kb, ke, halo, has_up, has_down):
        plan = []
        for d in range(3):
            f, kbase = self.field_planes(d)
            extra = 1 if d == 2 else 0
            o0, o1 = kb - kbase, ke - kbase
            up = (f[o1 - halo:o1], f[o1:o1 + halo + extra]) if has_up else None
            down = (f[o0:o0 + halo + extra], f[o0 - halo:o0]) if has_down else None
            plan.append((up, down))
        return plan
