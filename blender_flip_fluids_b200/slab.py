"""z-slab multi-GPU decomposition of the substep.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests). Rank g owns
the cell planes [g*K/n, (g+1)*K/n) of an I x J x K domain and stores ``halo`` more on each side.
Only nearest-neighbour exchanges are on the path, no all-reduce:

  1. ghost particles   particles within ``ghost`` cells of a slab face are copied to the
                       z-neighbour before P2G, so every owned face is summed wholly on its owner,
                       in the same (bin, global id) order as a single-GPU run;
  2. face halo         after P2G (where the reference's CPU projection would sit) the ``halo``
                       boundary planes of u, v, w are exchanged, then the field is saved;
  3. migration         after advection, particles whose cell left the slab move to the neighbour.

The driver is backend-agnostic: ``GpuBackend`` runs the stages in libffb200 on zero-copy tensor
views of the library's device buffers; the CPU tests plug in an oracle-backed backend to check
the exchange logic bit-for-bit against an undecomposed run.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

NSTREAM_FLIP = 6            # px py pz vx vy vz
NSTREAM_APIC = 15           # + 9 affine components


def slab_range(K: int, world: int, rank: int):
    """Owned cell planes of `rank`: contiguous, as equal as possible."""
    return (K * rank) // world, (K * (rank + 1)) // world


class _DevArray:
    """Adapter exposing a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _view(ptr, count, typestr, device):
    if not ptr or count == 0:
        return torch.empty(0, device=device, dtype={"<f4": torch.float32, "<u4": torch.int32, "|u1": torch.uint8}[typestr])
    if typestr == "<u4":                         # torch has no uint32 arithmetic: ids travel as int32 bit patterns
        typestr = "<i4"
    return torch.as_tensor(_DevArray(ptr, count, typestr), device=device)


def _staged(t):
    """gloo cannot move CUDA tensors point to point: stage them through host memory. (Used when two
    ranks share one GPU in the tests; on a multi-GPU box the backend is NCCL and nothing is staged.)"""
    return t.is_cuda and dist.get_backend() == "gloo"


def _sendrecv(items, wait=True):
    """One batched neighbour exchange. items: (send tensor or None, recv tensor or None, peer). wait=False (NCCL only)
    returns the outstanding requests: kernels launched meanwhile on the current stream overlap the transfer."""
    ops, staged = [], []
    for send, recv, peer in items:
        if send is not None:
            s = send.contiguous()
            ops.append(dist.P2POp(dist.isend, s.cpu() if _staged(s) else s, peer))
        if recv is not None:
            if _staged(recv):
                buf = torch.empty(recv.shape, dtype=recv.dtype)
                staged.append((recv, buf))
                ops.append(dist.P2POp(dist.irecv, buf, peer))
            else:
                ops.append(dist.P2POp(dist.irecv, recv, peer))
    if ops:
        reqs = dist.batch_isend_irecv(ops)
        if not wait and not staged:
            return reqs
        for r in reqs:
            r.wait()
    for recv, buf in staged:
        recv.copy_(buf)
    return []


def _all_gather(outs, t):
    if _staged(t):
        bufs = [torch.empty(o.shape, dtype=o.dtype) for o in outs]
        dist.all_gather(bufs, t.cpu())
        for o, b in zip(outs, bufs):
            o.copy_(b)
    else:
        dist.all_gather(outs, t)


class GpuBackend:
    """The stages on one GPU through the C ABI, with zero-copy views for the exchanges."""

    def __init__(self, I, J, K, dx, k_begin, k_end, halo, device_index, apic):
        from . import engine
        self.engine = engine
        self.device = torch.device("cuda", device_index)
        self.ctx = engine.FlipContext(I, J, K, dx, device=device_index, slab=(k_begin, k_end, halo))
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.apic = apic
        self.method = engine.APIC if apic else engine.FLIP
        self.I, self.J, self.K, self.dx = I, J, K, dx

    @property
    def nstream(self):
        return NSTREAM_APIC if self.apic else NSTREAM_FLIP

    def reserve(self, capacity):
        self.ctx.reserve_particles(capacity, self.apic)

    def particle_views(self, n=None):
        """(streams [nstream x n] as a list of 1-D views, ids int32 view) of the current buffer."""
        b = self.ctx.device_buffers()
        n = b.n if n is None else n
        ptrs = list(b.pos) + list(b.vel) + (list(b.aff) if self.apic else [])
        return [_view(p, n, "<f4", self.device) for p in ptrs], _view(b.ids, n, "<u4", self.device)

    def load_particles(self, streams, ids):
        n = int(ids.shape[0])
        b = self.ctx.device_buffers()
        if n > b.capacity:
            self.reserve(int(n * 1.25) + 1024)
        dst, dst_ids = self.particle_views(n)
        for d, s in zip(dst, streams):
            d.copy_(s)
        dst_ids.copy_(ids)
        self.ctx.set_num_particles(n, self.apic)

    def set_count(self, n):
        self.ctx.set_num_particles(n, self.apic)

    def field_planes(self, d, saved=False):
        b = self.ctx.device_buffers()
        ptr = (b.saved if saved else b.field)[d]
        plane = b.face_plane[d]
        return _view(ptr, b.face_count[d], "<f4", self.device).view(-1, plane), b.kbase

    def set_solid(self, phi, near):
        self.ctx.set_solid(phi, near)

    def p2g(self, radius):
        self.ctx.p2g(radius, self.method)

    def save_field(self):
        self.ctx.save_velocity_field()

    def g2p(self, ratio):
        self.ctx.g2p(self.method, ratio)

    def advect(self, dt, cfl=5.0, collide=True):
        self.ctx.advect(dt, cfl, collide)

    def synchronize(self):
        torch.cuda.synchronize(self.device)

    def set_window(self, k_lo, k_hi, mode):
        self.ctx.set_particle_window(k_lo, k_hi, mode)

    # ---- device-side plumbing (ffb200_slab.cu) ---------------------------------------------------
    fast = True

    def record_floats(self):
        return self.ctx.slab_record_floats()

    def new_block(self, capacity):
        """Packed-record buffer: capacity records + a 4-int header {count, overflowed, 0, 0}."""
        return torch.zeros(capacity * self.record_floats() + 4, dtype=torch.float32, device=self.device)

    def pack_layers(self, lo_a, hi_a, block_a, lo_b, hi_b, block_b, capacity):
        self.ctx.slab_pack_layers(lo_a, hi_a, block_a.data_ptr() if block_a is not None else 0, lo_b, hi_b,
                                  block_b.data_ptr() if block_b is not None else 0, capacity)

    def route(self, k_begin, k_end, block_up, block_down, capacity):
        return self.ctx.slab_route(k_begin, k_end, block_up.data_ptr() if block_up is not None else 0,
                                   block_down.data_ptr() if block_down is not None else 0, capacity)

    def route_begin(self, k_begin, k_end, block_up, block_down, capacity):
        self.ctx.slab_route_begin(k_begin, k_end, block_up.data_ptr() if block_up is not None else 0,
                                  block_down.data_ptr() if block_down is not None else 0, capacity)

    def route_end(self, leaving=None):
        """leaving: the header's count of particles leaving the resident set, if already known."""
        if leaving is None:
            return self.ctx.slab_route_end()
        return self.ctx.slab_route_end_known(leaving), None, None

    def new_block2(self, cap_m, cap_g):
        """Two-section packed buffer [cap_m migrants][cap_g ghost copies] + an 8-int header (ffb200_slab_route_ghosts_begin)."""
        return torch.zeros((cap_m + cap_g) * self.record_floats() + 8, dtype=torch.float32, device=self.device)

    def route_ghosts_begin(self, k_begin, k_end, ghost_layers, block_up, block_down, caps):
        self.ctx.slab_route_ghosts_begin(k_begin, k_end, ghost_layers, block_up.data_ptr() if block_up is not None else 0,
                                         block_down.data_ptr() if block_down is not None else 0, caps)

    def append_records(self, block, first_record, count, as_ghost=False):
        if count:
            self.ctx.slab_append(block.data_ptr() + 4 * first_record * self.record_floats(), count, as_ghost)

    def append(self, block, count, as_ghost=False):
        if count:
            self.ctx.slab_append(block.data_ptr(), count, as_ghost)

    def halo_plan(self, kb, ke, halo, has_up, has_down):
        """Cached (send, recv) plane views of u, v, w for the face-halo exchange (pointers of the
        face arrays never change). w stores one more plane; plane ke belongs to the upper slab."""
        plan = []
        for d in range(3):
            f, kbase = self.field_planes(d)
            extra = 1 if d == 2 else 0
            o0, o1 = kb - kbase, ke - kbase
            up = (f[o1 - halo:o1], f[o1:o1 + halo + extra]) if has_up else None
            down = (f[o0:o0 + halo + extra], f[o0 - halo:o0]) if has_down else None
            plan.append((up, down))
        return plan


class SlabSimulation:
    def __init__(self, I, J, K, dx, rank, world, backend, halo=7, ghost=2):
        self.I, self.J, self.K, self.dx = I, J, K, float(dx)
        self.rank, self.world = rank, world
        self.kb, self.ke = slab_range(K, world, rank)
        self.halo, self.ghost = halo, ghost
        self.backend = backend
        self.device = backend.device
        self.inv_dx = 1.0 / self.dx
        self.up = rank + 1 if rank + 1 < world else None
        self.down = rank - 1 if rank > 0 else None
        self.streams = None          # authoritative particle streams between substeps (list of 1-D tensors)
        self.ids = None
        self.exchanged_bytes = 0

    # ---- particles ----------------------------------------------------------------------------------
    def cell_k(self, pz):
        """floor(z * (1/dx)) in double: Grid3d::positionToGridIndex (grid3d.h:55-60)."""
        return torch.floor(pz.double() * self.inv_dx).to(torch.int64)

    def set_particles(self, streams, ids):
        self.streams = [s.contiguous() for s in streams]
        self.ids = ids.contiguous()

    def num_particles(self):
        return int(self.ids.shape[0])

    def _pack(self, mask):
        idx = torch.nonzero(mask, as_tuple=False).squeeze(1)
        rows = [s.index_select(0, idx) for s in self.streams]
        rows.append(self.ids.index_select(0, idx).view(torch.float32))
        return torch.stack(rows, 0) if idx.numel() else torch.empty((len(rows), 0), device=self.device)

    def _exchange(self, to_up, to_down):
        """Send one [rows, m] float32 block to each z-neighbour, receive theirs. Counts first."""
        rows = to_up.shape[0]
        dev = self.device
        cnt_out = {self.up: to_up, self.down: to_down}
        items, cnt_in = [], {}
        for peer, blk in cnt_out.items():
            if peer is None:
                continue
            cnt_in[peer] = torch.zeros(1, dtype=torch.int64, device=dev)
            items.append((torch.tensor([blk.shape[1]], dtype=torch.int64, device=dev), cnt_in[peer], peer))
        _sendrecv(items)
        items, recv = [], {}
        for peer, blk in cnt_out.items():
            if peer is None:
                continue
            m = int(cnt_in[peer].item())
            recv[peer] = torch.empty((rows, m), dtype=torch.float32, device=dev)
            if blk.shape[1]:
                self.exchanged_bytes += blk.numel() * 4
            items.append((blk if blk.shape[1] else None, recv[peer] if m else None, peer))
        _sendrecv(items)
        empty = torch.empty((rows, 0), dtype=torch.float32, device=dev)
        return recv.get(self.down, empty), recv.get(self.up, empty)

    def _append(self, blocks):
        ns = len(self.streams)
        for blk in blocks:
            if blk.shape[1] == 0:
                continue
            self.streams = [torch.cat([s, blk[i]]) for i, s in enumerate(self.streams)]
            self.ids = torch.cat([self.ids, blk[ns].view(torch.int32)])

    # ---- grids --------------------------------------------------------------------------------------
    def _halo_exchange(self):
        """Owned boundary planes of u, v, w -> the neighbours' halo planes (in place, zero copy)."""
        H = self.halo
        items = []
        for d in range(3):
            f, kbase = self.backend.field_planes(d)
            extra = 1 if d == 2 else 0                       # w has one more plane; plane ke belongs to the upper slab
            o0, o1 = self.kb - kbase, self.ke - kbase
            if self.up is not None:
                send = f[o1 - H:o1].contiguous()
                items.append((send, f[o1:o1 + H + extra], self.up))
                self.exchanged_bytes += send.numel() * 4
            if self.down is not None:
                send = f[o0:o0 + H + extra].contiguous()
                items.append((send, f[o0 - H:o0], self.down))
                self.exchanged_bytes += send.numel() * 4
        _sendrecv(items)

    # ---- one substep, device-side plumbing (GpuBackend) ----------------------------------------------
    INT_MIN, INT_MAX = -(2 ** 31), 2 ** 31 - 1

    def load_resident(self):
        """Make the backend's resident streams the authoritative particle state (fast path)."""
        self.backend.load_particles(self.streams, self.ids)
        self._resident = True
        self._ghosts_ready = False          # the resident set holds no ghost copies yet

    def _blocks(self, need):
        cap = getattr(self, "_block_cap", 0)
        if cap < need:
            cap = max(8192, int(need))
            self._block_cap = cap
            self._blk = {k: self.backend.new_block(cap) for k in ("up_send", "up_recv", "dn_send", "dn_recv")}
        return self._block_cap, self._blk

    def _swap_blocks(self, cap, b):
        """Exchange the fixed-size packed buffers with both neighbours. Returns (received from below,
        received from above, local overflow flag, headers [up_send, dn_send, up_recv, dn_recv])."""
        rows = self.backend.record_floats()
        items = []
        if self.up is not None:
            items.append((b["up_send"], b["up_recv"], self.up))
        if self.down is not None:
            items.append((b["dn_send"], b["dn_recv"], self.down))
        _sendrecv(items)
        hdr = torch.stack([b[k][cap * rows:].view(torch.int32) for k in ("up_send", "dn_send", "up_recv", "dn_recv")]).cpu()
        self.exchanged_bytes += (int(hdr[0, 0]) + int(hdr[1, 0])) * rows * 4
        overflow = bool(items) and int(hdr[:, 1].max()) != 0
        n_up = int(hdr[2, 0]) if self.up is not None else 0
        n_dn = int(hdr[3, 0]) if self.down is not None else 0
        return n_dn, n_up, overflow, hdr

    def _agree_max(self, value):
        """max over all ranks (buffer sizes of an exchange must match on both sides of every face)."""
        if self.world == 1:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64, device="cpu" if dist.get_backend() == "gloo" else self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def _tick(self, name):
        """FFB200_SLAB_PROFILE=1: wall-clock phase breakdown (synchronising; not for benchmarking)."""
        if not self._profile:
            return
        import time
        torch.cuda.synchronize(self.device)
        now = time.perf_counter()
        self._phase[name] = self._phase.get(name, 0.0) + now - self._t_last
        self._t_last = now

    def adopt_resident(self, n_owned):
        """The backend's resident streams were filled in place (benchscene.fill_dam_break): they are the authoritative
        particle state of this rank from now on."""
        self._resident = True
        self._ghosts_ready = False
        self._n_owned = int(n_owned)

    def step_fast(self, radius, ratio, dt, cfl=5.0, collide=True, apply_migration=True, projected_field=None, p2g_download=None,
                  overlap=False):
        """One substep with device-side plumbing. apply_migration=False is the fixed-batch
        benchmark mode (context in ffb200_set_fixed_batch): ghosts are added and removed and
        migrants are selected, packed and exchanged as usual, but the resident batch itself is
        left untouched so every step sees identical inputs."""
        import os, time
        if not hasattr(self, "_profile"):
            self._profile = os.environ.get("FFB200_SLAB_PROFILE") == "1"
            self._phase = {}
        if self._profile:
            torch.cuda.synchronize(self.device)
            self._t_last = time.perf_counter()
        be = self.backend
        g = self.ghost
        # 1. the very first substep after load_resident fetches its ghost layers with an exchange of its own;
        #    afterwards they arrive with the previous substep's migration (step 5)
        if not getattr(self, "_ghosts_ready", False):
            # buffer sizes and the retry decision are agreed by all ranks (rare path: one all-reduce each)
            n0 = be.ctx.n
            cap, b = self._blocks(max(8192, int(0.05 * self._agree_max(n0))))
            while True:
                be.pack_layers(self.ke - g, self.ke, b["up_send"] if self.up is not None else None,
                               self.kb, self.kb + g, b["dn_send"] if self.down is not None else None, cap)
                n_dn, n_up, overflow, hdr = self._swap_blocks(cap, b)
                if not self._agree_max(1 if overflow else 0):
                    break
                cap, b = self._blocks(cap * 2)
            be.append(b["dn_recv"], n_dn, as_ghost=True)
            be.append(b["up_recv"], n_up, as_ghost=True)
            self._ghosts_ready = True
            self._n_owned = n0
            # per face, the larger of what went either way: known to BOTH ranks of the face, so both size
            # the next exchange's buffers identically. No migrants yet: guess half a plane's worth.
            gu, gd = max(int(hdr[0, 0]), int(hdr[2, 0])), max(int(hdr[1, 0]), int(hdr[3, 0]))
            self._face_counts = {"up": (gu // (2 * g), gu), "dn": (gd // (2 * g), gd)}
        self._tick("ghosts")
        # 2. P2G on owned + ghost particles
        be.p2g(radius)
        self._tick("p2g")
        # 3. face halos, saved copy. overlap: the halo exchange is only STARTED here; the particles at least `halo` planes
        #    away from every face never sample a halo plane (CFL + stagger + trilinear < halo), so the first half of them
        #    is advanced while the planes travel (step 4a)
        inner = (self.kb + self.halo if self.down is not None else -(2 ** 30), self.ke - self.halo if self.up is not None else 2 ** 30)
        overlap = overlap and getattr(be, "set_window", None) is not None and self.world > 1
        mid = (max(inner[0], self.kb) + min(inner[1], self.ke)) // 2
        halo_pending = self._halo_exchange_fast(start_only=overlap)
        self._tick("halo")
        be.save_field()
        if p2g_download is not None:
            # e2e: the owned planes of the transferred field go to (pinned) host memory -- the CPU pressure solve's input
            for dst, src in zip(p2g_download, self._owned_field_views()):
                dst.copy_(src, non_blocking=True)
        if projected_field is not None:
            # the field the host's projection would hand back (stored planes incl. halo; device or pinned host tensors)
            for dst, src in zip(self._stored_field_views(), projected_field):
                dst.copy_(src.view(-1), non_blocking=True)
        self._tick("save+field")
        # 4. G2P + advection; the marked ghost copies ride along (a few %) and are dropped below.
        #    overlap: only the particles within `halo` planes of a slab face (and the ghost copies) can leave the slab or
        #    become ghost copies -- a substep moves a particle by at most CFL cells -- so they go first, the exchange of
        #    step 5 is started, and the interior particles are advanced while it is in flight.
        if overlap:
            be.set_window(inner[0], mid, 1)                   # 4a. interior, lower half: overlaps the halo exchange
            be.g2p(ratio)
            be.advect(dt, cfl, collide)
            self._tick("interior-A g2p+advect")
            for r in halo_pending:
                r.wait()
            self._halo_unpack(also_saved=True)
            if projected_field is not None:                   # the stand-in projection also covers the halo planes
                for dst, src in zip(self._stored_field_views(), projected_field):
                    dst.copy_(src.view(-1), non_blocking=True)
            self._tick("halo wait+unpack")
            be.set_window(inner[0], inner[1], 2)              # 4b. the particles near the faces + the ghost copies
        be.g2p(ratio)
        be.advect(dt, cfl, collide)
        self._tick("g2p+advect")
        # 5. ONE exchange: migrants + the ghost copies of the next substep. Ghosts of this substep are
        #    dropped; the end ranks keep whatever strayed past the domain
        kb = self.kb if self.down is not None else self.INT_MIN
        ke = self.ke if self.up is not None else self.INT_MAX
        up, dn = self.up is not None, self.down is not None
        caps, b2 = self._blocks2()
        faces, hdr = ("up", "dn"), None
        for attempt in range(3):
            if overlap and attempt > 0:
                be.set_window(inner[0], inner[1], 2)          # a repeated marking looks at the same particles
            be.route_ghosts_begin(kb, ke, g, b2["up_send"] if up else None, b2["dn_send"] if dn else None, caps)
            self._tick("route-mark")
            if overlap and attempt == 0:
                pending = self._swap_blocks2_start(b2, faces)
                be.set_window(mid, inner[1], 1)               # 4c. interior, upper half, while the blocks travel
                be.g2p(ratio)
                be.advect(dt, cfl, collide)
                be.set_window(0, 0, 0)
                self._tick("interior g2p+advect")
                for r in pending:
                    r.wait()
                h = self._swap_blocks2_headers(caps, b2)
            else:
                h = self._swap_blocks2(caps, b2, faces)   # one synchronisation serves the exchange and the routing
                if overlap:
                    be.set_window(0, 0, 0)
            self._tick("exchange")
            if hdr is None:
                hdr = h
            else:
                # rows 0, 1 (our send headers) always describe the LAST marking -- what route_end applies; rows 2, 3 (what
                # the neighbours sent) change only on the faces that were exchanged again
                hdr[0], hdr[1] = h[0], h[1]
                if "up" in faces:
                    hdr[2] = h[2]
                if "dn" in faces:
                    hdr[3] = h[3]
            # true counts of both directions per face (identical knowledge on both ranks of the face)
            self._face_counts = {"up": (max(int(hdr[0, 0]), int(hdr[2, 0])), max(int(hdr[0, 2]), int(hdr[2, 2]))),
                                 "dn": (max(int(hdr[1, 0]), int(hdr[3, 0])), max(int(hdr[1, 2]), int(hdr[3, 2])))}
            # a section that overflowed (migrants or ghost copies; both ranks of the face read it in the same two
            # headers) is exchanged again on that face alone, with capacities sized from the true counts the headers
            # carry: the marking is repeatable (the ghost bits of kept migrants are only applied by route_end)
            over = tuple(f for f, rows in (("up", (0, 2)), ("dn", (1, 3)))
                         if (up if f == "up" else dn) and any(int(hdr[r, 1]) or int(hdr[r, 3]) for r in rows))
            if not over or not apply_migration:
                break
            self.overflows = getattr(self, "overflows", 0) + 1
            faces = over
            caps, b2 = self._blocks2(resize=over)
        else:
            raise RuntimeError("slab exchange: sections still overflow after two resized exchanges (rank %d, capacities %s, "
                               "headers %s)" % (self.rank, caps, hdr.tolist()))
        n_owned, _, _ = be.route_end(int(hdr[0, 4]) if up else (int(hdr[1, 4]) if dn else None))
        mig_up, gh_up = (min(int(hdr[2, 0]), caps[0]), int(hdr[2, 2])) if up else (0, 0)
        mig_dn, gh_dn = (min(int(hdr[3, 0]), caps[2]), int(hdr[3, 2])) if dn else (0, 0)
        if apply_migration:
            be.append_records(b2["dn_recv"], 0, mig_dn)
            be.append_records(b2["up_recv"], 0, mig_up)
            n_owned += mig_dn + mig_up
        self._n_owned = n_owned
        be.append_records(b2["dn_recv"], caps[2], gh_dn, as_ghost=True)
        be.append_records(b2["up_recv"], caps[0], gh_up, as_ghost=True)
        self._tick("route-end+append")

    @staticmethod
    def _bucket(x):
        return (int(x) + 4095) // 4096 * 4096

    def _blocks2(self, resize=None):
        """Per-face section capacities (up migrants, up ghosts, down migrants, down ghosts) and buffers,
        derived from the previous exchange's counts on that face: twice the migrants (never below a quarter of a
        ghost layer: migrants arrive in bursts, a lattice layer reaches the face at once) + 4096, 1.15x the ghost
        copies + 8192, rounded up to 4096. Everything in a buffer travels, so the fit is kept tight. `resize` names the
        faces whose sections overflowed in this substep's exchange: their capacities are taken from the true counts
        (now in _face_counts), the other faces keep their buffers -- and what they received -- untouched."""
        (mu, gu), (md, gd) = self._face_counts["up"], self._face_counts["dn"]
        if resize is None and getattr(self, "force_initial_caps", None):
            caps = tuple(self.force_initial_caps)            # tests: undersized sections, so that every exchange is repeated
        elif resize is None:
            caps = (self._bucket(max(2 * mu, gu // 4) + 4096), self._bucket(1.15 * gu + 8192),
                    self._bucket(max(2 * md, gd // 4) + 4096), self._bucket(1.15 * gd + 8192))
        else:
            old = self._block2_caps
            caps = ((max(old[0], self._bucket(mu + 4096)), max(old[1], self._bucket(gu + 4096))) if "up" in resize else old[:2]) + \
                   ((max(old[2], self._bucket(md + 4096)), max(old[3], self._bucket(gd + 4096))) if "dn" in resize else old[2:])
        old_caps = getattr(self, "_block2_caps", None)
        if resize is None and old_caps is not None and not getattr(self, "force_initial_caps", None):
            # hysteresis: a buffer is kept while it is large enough and not more than twice what is needed (reallocating
            # -- and zero-filling -- 60 MB blocks because a count crossed a 4096 boundary costs more than the spare
            # bytes cost on NVLink)
            caps = tuple(o if n <= o <= 2 * n else n for o, n in zip(old_caps, caps))
        if old_caps != caps:
            nb = self.backend.new_block2
            blk = dict(getattr(self, "_blk2", {}))
            if old_caps is None or old_caps[:2] != caps[:2]:
                blk["up_send"], blk["up_recv"] = nb(caps[0], caps[1]), nb(caps[0], caps[1])
            if old_caps is None or old_caps[2:] != caps[2:]:
                blk["dn_send"], blk["dn_recv"] = nb(caps[2], caps[3]), nb(caps[2], caps[3])
            self._block2_caps, self._blk2 = caps, blk
        return caps, self._blk2

    def _swap_items(self, b, faces):
        items = []
        if self.up is not None and "up" in faces:
            items.append((b["up_send"], b["up_recv"], self.up))
        if self.down is not None and "dn" in faces:
            items.append((b["dn_send"], b["dn_recv"], self.down))
        return items

    def _swap_blocks2_start(self, b, faces=("up", "dn")):
        """Start the exchange of the two-section buffers without waiting (NCCL; gloo staging completes it at once)."""
        return _sendrecv(self._swap_items(b, faces), wait=False)

    def _swap_blocks2(self, caps, b, faces=("up", "dn")):
        """Exchange the two-section buffers with the neighbours on `faces`; returns the four 8-int headers
        [up_send, dn_send, up_recv, dn_recv] (rows of absent neighbours are zero)."""
        _sendrecv(self._swap_items(b, faces))
        return self._swap_blocks2_headers(caps, b)

    def _swap_blocks2_headers(self, caps, b):
        rows = self.backend.record_floats()
        off = {"up_send": (caps[0] + caps[1]) * rows, "up_recv": (caps[0] + caps[1]) * rows,
               "dn_send": (caps[2] + caps[3]) * rows, "dn_recv": (caps[2] + caps[3]) * rows}
        hdr = torch.stack([b[k][off[k]:].view(torch.int32) for k in ("up_send", "dn_send", "up_recv", "dn_recv")]).cpu()
        if self.up is None:
            hdr[0].zero_(); hdr[2].zero_()
        if self.down is None:
            hdr[1].zero_(); hdr[3].zero_()
        self.exchanged_bytes += int(hdr[0, 0] + hdr[0, 2] + hdr[1, 0] + hdr[1, 2]) * rows * 4
        return hdr

    def _stored_field_views(self):
        v = getattr(self, "_stored_views", None)
        if v is None:
            b = self.backend.ctx.device_buffers()
            v = self._stored_views = [_view(b.field[d], b.face_count[d], "<f4", self.device) for d in range(3)]
        return v

    def _owned_field_views(self):
        v = getattr(self, "_owned_views", None)
        if v is None:
            v = []
            for d in range(3):
                f, kbase = self.backend.field_planes(d)
                v.append(f[self.kb - kbase:self.ke - kbase + (1 if d == 2 and self.up is None else 0)])
            self._owned_views = v
        return v

    def _halo_exchange_fast(self, start_only=False):
        """Face halos of u, v, w: ONE message per face and direction (the three fields' plane ranges packed into one
        buffer; 12 separate NCCL point-to-point operations per middle rank cost more in latency than the 21 MB per
        face cost in bandwidth)."""
        plan = getattr(self, "_halo_plan", None)
        if plan is None:
            plan = self.backend.halo_plan(self.kb, self.ke, self.halo, self.up is not None, self.down is not None)
            packed = {}
            for side, peer in (("up", self.up), ("dn", self.down)):
                if peer is None:
                    continue
                pairs = [p[0 if side == "up" else 1] for p in plan]              # (send view, recv view) per field
                ns = [v[0].numel() for v in pairs]
                nr = [v[1].numel() for v in pairs]
                packed[side] = (peer, pairs, ns, nr, torch.empty(sum(ns), dtype=torch.float32, device=self.device),
                                torch.empty(sum(nr), dtype=torch.float32, device=self.device))
            plan = self._halo_plan = packed
        items = []
        for side, (peer, pairs, ns, nr, sbuf, rbuf) in plan.items():
            o = 0
            for (send, _), m in zip(pairs, ns):
                sbuf[o:o + m].copy_(send.reshape(-1))
                o += m
            items.append((sbuf, rbuf, peer))
            self.exchanged_bytes += sbuf.numel() * 4
        pending = _sendrecv(items, wait=not start_only)
        if start_only:
            return pending
        self._halo_unpack()
        return []

    def _halo_unpack(self, also_saved=False):
        """Received halo planes -> the halo planes of u, v, w (and of the saved copy, when the copy was taken before
        they arrived)."""
        for side, (peer, pairs, ns, nr, sbuf, rbuf) in self._halo_plan.items():
            o = 0
            for d, ((_, recv), m) in enumerate(zip(pairs, nr)):
                recv.copy_(rbuf[o:o + m].view(recv.shape))
                o += m
        if also_saved:
            for d in range(3):
                f, kbase = self.backend.field_planes(d)
                sv, _ = self.backend.field_planes(d, saved=True)
                extra = 1 if d == 2 else 0
                o0, o1 = self.kb - kbase, self.ke - kbase
                if self.up is not None:
                    sv[o1:o1 + self.halo + extra].copy_(f[o1:o1 + self.halo + extra])
                if self.down is not None:
                    sv[o0 - self.halo:o0].copy_(f[o0 - self.halo:o0])

    def sync_from_backend(self):
        """Pull the resident streams back into self.streams / self.ids (tests, gather)."""
        s, ids = self.backend.particle_views()
        own = torch.nonzero(ids.view(torch.int32) >= 0, as_tuple=False).squeeze(1)    # ghost copies carry the top id bit
        self.streams = [t.index_select(0, own) for t in s]
        self.ids = ids.index_select(0, own)

    # ---- one substep, generic plumbing (any backend; used by the CPU tests) ------------------------------
    def step(self, radius, ratio, dt, cfl=5.0, collide=True):
        kz = self.cell_k(self.streams[2])
        # 1. ghost particles for P2G
        g = self.ghost
        up_mask = (kz >= self.ke - g) if self.up is not None else torch.zeros_like(kz, dtype=torch.bool)
        dn_mask = (kz < self.kb + g) if self.down is not None else torch.zeros_like(kz, dtype=torch.bool)
        from_down, from_up = self._exchange(self._pack(up_mask), self._pack(dn_mask))
        ns = len(self.streams)
        streams = [torch.cat([s, from_down[i], from_up[i]]) for i, s in enumerate(self.streams)]
        ids = torch.cat([self.ids, from_down[ns].view(torch.int32), from_up[ns].view(torch.int32)])
        self.backend.load_particles(streams, ids)
        # 2. P2G on owned + ghost particles (sorts; owned faces are complete)
        self.backend.p2g(radius)
        sorted_streams, sorted_ids = self.backend.particle_views()
        ghost_mask = ~self._owned(self.cell_k(sorted_streams[2]))
        # 3. halo planes, then the saved copy (the CPU projection would run between the two)
        self._halo_exchange()
        self.backend.save_field()
        # 4. G2P + advection (in place on the sorted streams; ghosts ride along and are dropped below)
        self.backend.g2p(ratio)
        self.backend.advect(dt, cfl, collide)
        sorted_streams, sorted_ids = self.backend.particle_views()
        # 5. migration
        knew = self.cell_k(sorted_streams[2])
        alive = ~ghost_mask
        stay = alive & self._owned(knew)
        go_up = alive & (knew >= self.ke) if self.up is not None else torch.zeros_like(stay)
        go_dn = alive & (knew < self.kb) if self.down is not None else torch.zeros_like(stay)
        if self.up is None:
            stay = stay | (alive & (knew >= self.ke))           # top rank keeps what left the domain upwards
        if self.down is None:
            stay = stay | (alive & (knew < self.kb))
        self.streams, self.ids = sorted_streams, sorted_ids
        up_blk, dn_blk = self._pack(go_up), self._pack(go_dn)
        idx = torch.nonzero(stay, as_tuple=False).squeeze(1)
        self.streams = [s.index_select(0, idx) for s in sorted_streams]
        self.ids = sorted_ids.index_select(0, idx)
        from_down, from_up = self._exchange(up_blk, dn_blk)
        self._append([from_down, from_up])

    def _owned(self, kz):
        return (kz >= self.kb) & (kz < self.ke)

    # ---- helpers for tests / bench ----------------------------------------------------------------------
    def gather_particles(self):
        """All particles on every rank, ordered by global id (tests)."""
        ns = len(self.streams)
        blk = torch.stack(self.streams + [self.ids.view(torch.float32)], 0).contiguous()
        counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        _all_gather(counts, torch.tensor([blk.shape[1]], dtype=torch.int64, device=self.device))
        parts = [torch.empty((ns + 1, int(c.item())), dtype=torch.float32, device=self.device) for c in counts]
        mx = max(int(c.item()) for c in counts)
        padded = torch.zeros((ns + 1, mx), dtype=torch.float32, device=self.device)
        padded[:, :blk.shape[1]] = blk
        bufs = [torch.zeros_like(padded) for _ in range(self.world)]
        _all_gather(bufs, padded)
        for r in range(self.world):
            parts[r] = bufs[r][:, :int(counts[r].item())]
        allp = torch.cat(parts, 1)
        ids = allp[ns].contiguous().view(torch.int32).to(torch.int64) & 0xffffffff
        order = torch.argsort(ids)
        return allp[:ns, order], ids[order]


def make_slab_dam_break(I, J, K, dx, rank, world, ppc, v0, apic, seed, device):
    """This rank's particles of the dam break I x J x K (full-z columns), generated
    independently per rank (seed + rank); global ids are offset by the lower ranks' counts."""
    from . import scenes
    kb, ke = slab_range(K, world, rank)
    rng = np.random.default_rng(seed + 7919 * rank)
    k0, k1 = max(3, kb), min(K - 3, ke)
    pos = scenes._seed_cells(3, max(4, int(0.4 * I)), 3, max(4, int(0.8 * J)), k0, k1, dx, ppc, rng)
    n = pos.shape[0]
    vel = (rng.uniform(-1.0, 1.0, size=(n, 3)) * v0).astype(np.float32)
    cols = [pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2]]
    if apic:
        aff = (rng.uniform(-1.0, 1.0, size=(n, 9)) * (0.1 / dx)).astype(np.float32)
        cols += [aff[:, q] for q in range(9)]
    streams = [torch.from_numpy(np.ascontiguousarray(c)).to(device) for c in cols]
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    if world > 1:
        _all_gather(counts, torch.tensor([n], dtype=torch.int64, device=device))
    else:
        counts[0][0] = n
    base = int(sum(int(c.item()) for c in counts[:rank]))
    ids = (torch.arange(n, dtype=torch.int64, device=device) + base).to(torch.int32)
    return streams, ids, int(sum(int(c.item()) for c in counts))
