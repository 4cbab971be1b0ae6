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

__global__ void k_route_mark(Streams cur, Streams dat, int n, double inv_dx, int k_begin, int k_end, float *block_up,
                             float *block_down, int cap, int remove_migrants, int *counters, uint32_t *holes) {
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
}

// holes below the new count, and survivors at or above it
__global__ void k_fill_collect(Streams cur, Streams dat, int n, int n_new, int nholes, double inv_dx, int k_begin, int k_end,
                               int has_up, int has_down, int remove_migrants, const uint32_t *holes, int *counters,
                               uint32_t *front_hole, uint32_t *tail_stay) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_front = false;
    uint32_t h = 0;
    if (t < nholes) {
        h = holes[t];
        is_front = h < (uint32_t)n_new;
    }
    bool is_tail_stay = false;
    const int j = n_new + t;
    if (t < nholes && j < n) {
        const int k = __double2int_rd((double)dat.s[2][j] * inv_dx);
        const bool ghost = (cur.ids[j] & kGhostBit) != 0u;
        const bool up = !ghost && k >= k_end && has_up;
        const bool down = !ghost && k < k_begin && has_down;
        is_tail_stay = !(ghost || (remove_migrants && (up || down)));
    }
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
    int k_begin, k_end, has_up, has_down, n;
};
static RouteState g_route;      // one routing in flight per process (one context per rank)

int launch_route_begin(Context &c, int k_begin, int k_end, float *block_up, float *block_down, int cap) {
    int launches = 0;
    const int n = c.n;
    const int rows = slab_rows(c);
    const bool fixed = c.nondestructive;
    Streams cur = streams_of(c, c.cur);
    Streams dat = fixed ? streams_of(c, c.cur ^ 1) : cur;      // fixed batch: advect / G2P wrote to the spare buffer
    uint32_t *holes = c.sort.key[1];                           // key/val scratch is free after P2G
    FFB_CUDA(cudaMemsetAsync(c.slab_counters, 0, 4 * sizeof(int), c.stream));
    if (n > 0) {
        k_route_mark<<<(n + 255) / 256, 256, 0, c.stream>>>(cur, dat, n, c.g.inv_dx, k_begin, k_end, block_up, block_down, cap,
                                                            fixed ? 0 : 1, c.slab_counters, holes);
        launches++;
    }
    if (block_up) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 1, cap, rows, block_up); launches++; }
    if (block_down) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 2, cap, rows, block_down); launches++; }
    g_route = RouteState{k_begin, k_end, block_up != nullptr, block_down != nullptr, n};
    FFB_CUDA(cudaGetLastError());
    return launches;
}

int launch_route_end(Context &c, int counts_host[3]) {
    int launches = 0;
    const int n = g_route.n, k_begin = g_route.k_begin, k_end = g_route.k_end;
    const bool fixed = c.nondestructive;
    Streams cur = streams_of(c, c.cur);
    Streams dat = fixed ? streams_of(c, c.cur ^ 1) : cur;
    uint32_t *holes = c.sort.key[1], *front_hole = c.sort.val[1], *tail_stay = c.sort.key[0];
    int h[4] = {0, 0, 0, 0};
    FFB_CUDA(cudaMemcpyAsync(h, c.slab_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    FFB_CUDA(cudaStreamSynchronize(c.stream));
    const int nholes = h[3], n_new = n - nholes;
    if (nholes > 0 && n_new > 0) {
        FFB_CUDA(cudaMemsetAsync(c.slab_counters, 0, 2 * sizeof(int), c.stream));
        k_fill_collect<<<(nholes + 255) / 256, 256, 0, c.stream>>>(cur, dat, n, n_new, nholes, c.g.inv_dx, k_begin, k_end,
                                                                   g_route.has_up, g_route.has_down, fixed ? 0 : 1,
                                                                   holes, c.slab_counters, front_hole, tail_stay);
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
