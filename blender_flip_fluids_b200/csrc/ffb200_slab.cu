// ffb200_slab.cu -- device-side plumbing of the z-slab decomposition:
// selecting ghost layers, dropping ghosts, splitting migrants from stayers, appending received
// particles. Everything works on the resident SoA streams; particles travel between ranks as
// packed records [count][rows] (rows = 6 or 15 attribute floats + the global id) in caller-owned
// device buffers whose last 4 ints are a header {count, overflow, 0, 0} written on the device, so
// a fixed-capacity buffer can be sent over NCCL without a host round trip for the size.
//
// Slot allocation uses warp-aggregated integer atomics; the order inside a packed buffer or a
// compacted stream is therefore arbitrary, which is harmless: the next substep re-sorts by
// (bin, global id), so every downstream result is deterministic.
#include "ffb200_ctx.h"

namespace ffb200 {

namespace {

struct Streams {
    float *s[15];
    uint32_t *ids;
    int ns;          // 6 or 15
};

__device__ __forceinline__ int warp_slot(bool pred, int *counter) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!pred) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void write_record(const Streams &src, int j, float *block, int cap, int slot) {
    if (slot >= cap) return;                                  // overflow is reported through the header
    float *r = block + (size_t)slot * (src.ns + 1);
    float v[15];
#pragma unroll
    for (int t = 0; t < 15; t++)
        if (t < src.ns) v[t] = src.s[t][j];
    const uint32_t id = src.ids[j];
#pragma unroll
    for (int t = 0; t < 15; t++)
        if (t < src.ns) r[t] = v[t];
    r[src.ns] = __uint_as_float(id);
}

// Copy (not move) the particles of two cell-plane ranges into two packed buffers.
__global__ void k_pack_layers(Streams src, int n, double inv_dx, int lo_a, int hi_a, float *block_a, int lo_b, int hi_b,
                              float *block_b, int cap, int *counters) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0x7fffffff;
    if (j < n) k = __double2int_rd((double)src.s[2][j] * inv_dx);
    const bool in_a = j < n && block_a && k >= lo_a && k < hi_a;
    const bool in_b = j < n && block_b && k >= lo_b && k < hi_b;
    const int sa = warp_slot(in_a, counters + 0);
    const int sb = warp_slot(in_b, counters + 1);
    if (in_a) write_record(src, j, block_a, cap, sa);
    if (in_b) write_record(src, j, block_b, cap, sb);
}

// Route every particle by its cell plane: [k_begin, k_end) stays, above goes to block_up, below
// to block_down (a null block keeps those particles). Particles whose id carries kGhostBit (set
// by k_append for ghost copies) always leave. A bin never mixes ghosts and owned particles (the
// slab boundary is a cell plane), so the mark does not disturb the (bin, id) order of the sort.
//
// Leavers are removed by HOLE FILLING instead of compacting every stream: pass 1 packs the
// migrants and lists the holes; pass 2 pairs the holes in the surviving prefix [0, n - L) with
// the stayers of the tail [n - L, n); pass 3 moves just those. Traffic is proportional to the
// few percent that leave, not to n.
//
// `dec`/`dat` may differ from `cur` in fixed-batch mode, where advection and G2P wrote to the
// spare buffer: the decision and the packed payload come from there, the resident batch is
// left as it is (migrants are packed and sent, but not removed).
constexpr uint32_t kGhostBit = 0x80000000u;

struct FaceCaps {
    int up_m, up_g, dn_m, dn_g;      // record capacities of the migrant / ghost sections of the up and down blocks
};

__global__ void k_route_mark(Streams cur, Streams dat, int n, double inv_dx, int k_begin, int k_end, float *block_up,
                             float *block_down, int cap, int remove_migrants, int *counters, uint32_t *holes,
                             uint8_t *leave_flag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool ghost = false, up = false, down = false;
    if (j < n) {
        const int k = __double2int_rd((double)dat.s[2][j] * inv_dx);
        ghost = (cur.ids[j] & kGhostBit) != 0u;
        up = !ghost && k >= k_end && block_up;
        down = !ghost && k < k_begin && block_down;
    }
    const bool leave = ghost || (remove_migrants && (up || down));
    const int su = warp_slot(up, counters + 1);
    const int sd = warp_slot(down, counters + 2);
    const int sh = warp_slot(leave, counters + 3);
    if (up || down) {
        Streams rec = dat;
        rec.ids = cur.ids;
        write_record(rec, j, up ? block_up : block_down, cap, up ? su : sd);
    }
    if (leave) holes[sh] = (uint32_t)j;
    if (j < n) leave_flag[j] = leave ? 1 : 0;
}

// The same routing with the NEXT substep's ghost exchange folded in, so that one neighbour
// exchange per substep carries both. Blocks have two sections, [migrants][ghost copies] (their
// capacities are per face: FaceCaps), and an 8-int header {migrants, overflow, ghosts, overflow, holes, 0...}.
//  * an owned particle that stays and sits within `g` cell planes of a slab face is copied into the
//    ghost section for that neighbour (it is what k_pack_layers would select at the start of the
//    next substep);
//  * a migrant that lands within `g` planes beyond the face is sent AND kept here with the ghost
//    bit set: its new owner would send it straight back as a ghost copy otherwise;
//  * ghost copies of the substep that just ended leave, as before.
// Fixed-batch mode (remove_migrants = 0): nothing is removed or converted; ghost copies are taken
// from the resident (pristine) batch, exactly what a start-of-substep exchange would send.
__global__ void k_route_mark_ghosts(Streams cur, Streams dat, int n, double inv_dx, int k_begin, int k_end, int g,
                                    float *block_up, float *block_down, FaceCaps caps, int remove_migrants, int *counters,
                                    uint32_t *holes, uint8_t *leave_flag, Window win) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool ghost = false, up = false, down = false, gup = false, gdown = false, keep = false;
    // with a window only the particles near the slab faces (and the ghost copies) are looked at: the others can neither
    // leave nor be ghost copies (their flags were cleared by the launcher). Whole warps drop out together almost always.
    const bool look = j < n && !window_skip(win, j);
    if (look) {
        const int kd = __double2int_rd((double)dat.s[2][j] * inv_dx);      // where the particle is going
        const int kc = __double2int_rd((double)cur.s[2][j] * inv_dx);      // where the resident copy is
        ghost = (cur.ids[j] & kGhostBit) != 0u;
        up = !ghost && kd >= k_end && block_up;
        down = !ghost && kd < k_begin && block_down;
        const bool owned_after = !ghost && !(remove_migrants && (up || down));
        gup = owned_after && block_up && kc >= k_end - g;
        gdown = owned_after && block_down && kc < k_begin + g;
        keep = remove_migrants && ((up && kd < k_end + g) || (down && kd >= k_begin - g));
    }
    const int su = warp_slot(up, counters + 1);
    const int sd = warp_slot(down, counters + 2);
    // a migrant that does not fit the buffer is not lost: it stays here (outside its slab for one
    // substep; the header reports the overflow and the host enlarges the buffers)
    const bool sent = (up && su < caps.up_m) || (down && sd < caps.dn_m);
    keep = keep && sent;
    const bool leave = ghost || (remove_migrants && sent && !keep);
    const int sh = warp_slot(leave, counters + 3);
    const int sgu = warp_slot(gup, counters + 4);
    const int sgd = warp_slot(gdown, counters + 5);
    const int rows = cur.ns + 1;
    if (sent) {
        Streams rec = dat;
        rec.ids = cur.ids;
        write_record(rec, j, up ? block_up : block_down, up ? caps.up_m : caps.dn_m, up ? su : sd);
    }
    if (gup) write_record(cur, j, block_up + (size_t)caps.up_m * rows, caps.up_g, sgu);
    if (gdown) write_record(cur, j, block_down + (size_t)caps.dn_m * rows, caps.dn_g, sgd);
    // `keep` (sent, and kept here as the new owner's ghost copy) is only FLAGGED: the ghost bit is set by
    // launch_route_end, so that the marking can be repeated with larger capacities when a section overflowed
    if (leave) holes[sh] = (uint32_t)j;
    if (look) leave_flag[j] = leave ? 1 : (keep ? 2 : 0);
}

__global__ void k_apply_keep(uint32_t *__restrict__ ids, const uint8_t *__restrict__ leave_flag, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && leave_flag[j] == 2) ids[j] |= kGhostBit;
}

__global__ void k_write_header2(const int *counters, int which_m, int which_g, int cap, int cap_g, int rows, float *block) {
    int *h = reinterpret_cast<int *>(block + ((size_t)cap + cap_g) * rows);
    const int m = counters[which_m], gc = counters[which_g];
    h[0] = m; h[1] = m > cap ? 1 : 0;
    h[2] = gc; h[3] = gc > cap_g ? 1 : 0;
    h[4] = counters[3];                                        // particles leaving the resident set (holes to fill)
    h[5] = h[6] = h[7] = 0;
}

// holes below the new count, and survivors at or above it
__global__ void k_fill_collect(int n, int n_new, int nholes, const uint8_t *leave_flag, const uint32_t *holes, int *counters,
                               uint32_t *front_hole, uint32_t *tail_stay) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_front = false;
    uint32_t h = 0;
    if (t < nholes) {
        h = holes[t];
        is_front = h < (uint32_t)n_new;
    }
    const int j = n_new + t;
    const bool is_tail_stay = t < nholes && j < n && leave_flag[j] != 1;
    const int sa = warp_slot(is_front, counters + 0);
    const int sb = warp_slot(is_tail_stay, counters + 1);
    if (is_front) front_hole[sa] = h;
    if (is_tail_stay) tail_stay[sb] = (uint32_t)j;
}

__global__ void k_fill_move(Streams cur, const int *counters, const uint32_t *front_hole, const uint32_t *tail_stay) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= counters[0]) return;                             // == counters[1]
    const uint32_t dst = front_hole[t], src = tail_stay[t];
#pragma unroll
    for (int q = 0; q < 15; q++)
        if (q < cur.ns) cur.s[q][dst] = cur.s[q][src];
    cur.ids[dst] = cur.ids[src];
}

__global__ void k_write_header(const int *counters, int which, int cap, int rows, float *block) {
    int *h = reinterpret_cast<int *>(block + (size_t)cap * rows);
    const int c = counters[which];
    h[0] = c;
    h[1] = c > cap ? 1 : 0;
    h[2] = 0;
    h[3] = 0;
}

__global__ void k_append(Streams dst, int offset, const float *__restrict__ block, int count, uint32_t id_or) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const float *r = block + (size_t)j * (dst.ns + 1);
#pragma unroll
    for (int t = 0; t < 15; t++)
        if (t < dst.ns) dst.s[t][offset + j] = r[t];
    dst.ids[offset + j] = __float_as_uint(r[dst.ns]) | id_or;
}

Streams streams_of(Context &c, int buf) {
    Streams s;
    ParticleSoA &p = c.soa[buf];
    for (int q = 0; q < 3; q++) { s.s[q] = p.p[q]; s.s[3 + q] = p.v[q]; }
    for (int q = 0; q < 9; q++) s.s[6 + q] = c.has_affine ? p.a[q] : nullptr;
    s.ids = p.orig;
    s.ns = c.has_affine ? 15 : 6;
    return s;
}

}  // namespace

int slab_rows(Context &c) { return (c.has_affine ? 15 : 6) + 1; }

int launch_pack_layers(Context &c, int lo_a, int hi_a, float *block_a, int lo_b, int hi_b, float *block_b, int cap) {
    int launches = 0;
    FFB_CUDA(cudaMemsetAsync(c.slab_counters, 0, 4 * sizeof(int), c.stream));
    if (c.n > 0 && (block_a || block_b)) {
        k_pack_layers<<<(c.n + 255) / 256, 256, 0, c.stream>>>(streams_of(c, c.cur), c.n, c.g.inv_dx, lo_a, hi_a, block_a,
                                                               lo_b, hi_b, block_b, cap, c.slab_counters);
        launches++;
    }
    const int rows = slab_rows(c);
    if (block_a) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 0, cap, rows, block_a); launches++; }
    if (block_b) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 1, cap, rows, block_b); launches++; }
    FFB_CUDA(cudaGetLastError());
    return launches;
}

// Routing is split in two so the host can overlap the neighbour exchange with it:
// launch_route_begin packs the migrants and lists the holes (asynchronous); launch_route_end
// reads the counts (synchronises the stream), fills the holes and sets the new particle count.
struct RouteState {
    int k_begin, k_end, has_up, has_down, n, ghosts;
};
static RouteState g_route;      // one routing in flight per process (one context per rank)

int launch_route_begin(Context &c, int k_begin, int k_end, float *block_up, float *block_down, int cap, int ghost_layers,
                       const int *caps) {
    int launches = 0;
    const int n = c.n;
    const int rows = slab_rows(c);
    const bool fixed = c.nondestructive;
    Streams cur = streams_of(c, c.cur);
    Streams dat = fixed ? streams_of(c, c.cur ^ 1) : cur;      // fixed batch: advect / G2P wrote to the spare buffer
    uint32_t *holes = c.sort.key[1];                           // key/val scratch is free after P2G
    uint8_t *leave_flag = reinterpret_cast<uint8_t *>(c.sort.val[0]);
    FFB_CUDA(cudaMemsetAsync(c.slab_counters, 0, 8 * sizeof(int), c.stream));
    if (ghost_layers > 0) {
        if (n > 0) {
            if (c.window.mode != 0) FFB_CUDA(cudaMemsetAsync(leave_flag, 0, (size_t)n, c.stream));
            k_route_mark_ghosts<<<(n + 255) / 256, 256, 0, c.stream>>>(cur, dat, n, c.g.inv_dx, k_begin, k_end, ghost_layers,
                                                                       block_up, block_down, FaceCaps{caps[0], caps[1], caps[2], caps[3]},
                                                                       fixed ? 0 : 1, c.slab_counters, holes, leave_flag, c.window);
            launches++;
        }
        if (block_up) { k_write_header2<<<1, 1, 0, c.stream>>>(c.slab_counters, 1, 4, caps[0], caps[1], rows, block_up); launches++; }
        if (block_down) { k_write_header2<<<1, 1, 0, c.stream>>>(c.slab_counters, 2, 5, caps[2], caps[3], rows, block_down); launches++; }
    } else {
        if (n > 0) {
            k_route_mark<<<(n + 255) / 256, 256, 0, c.stream>>>(cur, dat, n, c.g.inv_dx, k_begin, k_end, block_up, block_down,
                                                                cap, fixed ? 0 : 1, c.slab_counters, holes, leave_flag);
            launches++;
        }
        if (block_up) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 1, cap, rows, block_up); launches++; }
        if (block_down) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 2, cap, rows, block_down); launches++; }
    }
    g_route = RouteState{k_begin, k_end, block_up != nullptr, block_down != nullptr, n, ghost_layers > 0 ? 1 : 0};
    FFB_CUDA(cudaGetLastError());
    return launches;
}

// known_holes >= 0: the caller already read the number of leaving particles from a block header (h[4] of
// ffb200_slab_route_ghosts_begin's blocks, available after its exchange): no second synchronisation.
int launch_route_end(Context &c, int counts_host[3], int known_holes) {
    int launches = 0;
    const int n = g_route.n;
    uint32_t *holes = c.sort.key[1], *front_hole = c.sort.val[1], *tail_stay = c.sort.key[0];
    const uint8_t *leave_flag = reinterpret_cast<const uint8_t *>(c.sort.val[0]);
    Streams cur = streams_of(c, c.cur);
    int h[4] = {0, 0, 0, known_holes};
    if (known_holes < 0) {
        FFB_CUDA(cudaMemcpyAsync(h, c.slab_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaStreamSynchronize(c.stream));
    }
    const int nholes = h[3], n_new = n - nholes;
    if (g_route.ghosts && n > 0) {                             // migrants kept as ghost copies get their ghost bit now
        k_apply_keep<<<(n + 255) / 256, 256, 0, c.stream>>>(cur.ids, leave_flag, n);
        launches++;
    }
    if (nholes > 0 && n_new > 0) {
        FFB_CUDA(cudaMemsetAsync(c.slab_counters, 0, 2 * sizeof(int), c.stream));
        k_fill_collect<<<(nholes + 255) / 256, 256, 0, c.stream>>>(n, n_new, nholes, leave_flag, holes, c.slab_counters,
                                                                   front_hole, tail_stay);
        // the two lists have the same length by construction (holes in the prefix == survivors in the tail)
        k_fill_move<<<(nholes + 255) / 256, 256, 0, c.stream>>>(cur, c.slab_counters, front_hole, tail_stay);
        launches += 2;
    }
    counts_host[0] = n_new;
    counts_host[1] = h[1];
    counts_host[2] = h[2];
    c.n = n_new;
    c.sorted = false;
    FFB_CUDA(cudaGetLastError());
    return launches;
}

int launch_append(Context &c, const float *block, int count, bool as_ghost) {
    if (count <= 0) return 0;
    k_append<<<(count + 255) / 256, 256, 0, c.stream>>>(streams_of(c, c.cur), c.n, block, count, as_ghost ? kGhostBit : 0u);
    FFB_CUDA(cudaGetLastError());
    c.n += count;
    c.sorted = false;
    return 1;
}

}  // namespace ffb200
