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
#pragma unroll
    for (int t = 0; t < 15; t++)
        if (t < src.ns) r[t] = src.s[t][j];
    r[src.ns] = __uint_as_float(src.ids[j]);
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

// Route every particle by its cell plane: [k_begin, k_end) stays (compacted into dst), above goes
// to block_up, below to block_down; a null block drops those particles.
//
// Particles whose id carries kGhostBit (set by k_append for ghost copies) are dropped here
// whatever their position. A bin never mixes ghosts and owned particles (the slab boundary is a
// cell plane), so the mark does not disturb the (bin, id) order of the sort.
constexpr uint32_t kGhostBit = 0x80000000u;

__global__ void k_route(Streams src, Streams dst, int n, double inv_dx, int k_begin, int k_end, float *block_up,
                        float *block_down, int cap, int *counters) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0;
    bool alive = false;
    if (j < n) {
        k = __double2int_rd((double)src.s[2][j] * inv_dx);
        alive = (src.ids[j] & kGhostBit) == 0u;
    }
    const bool stay = alive && k >= k_begin && k < k_end;
    const bool up = alive && k >= k_end && block_up;
    const bool down = alive && k < k_begin && block_down;
    const int ss = warp_slot(stay, counters + 0);
    const int su = warp_slot(up, counters + 1);
    const int sd = warp_slot(down, counters + 2);
    if (stay) {
#pragma unroll
        for (int t = 0; t < 15; t++)
            if (t < src.ns) dst.s[t][ss] = src.s[t][j];
        dst.ids[ss] = src.ids[j];
    }
    if (up) write_record(src, j, block_up, cap, su);
    if (down) write_record(src, j, block_down, cap, sd);
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

// Returns the launch count; counts_host[3] = {stay, up, down} after a stream synchronisation.
int launch_route(Context &c, int k_begin, int k_end, float *block_up, float *block_down, int cap, int counts_host[3]) {
    int launches = 0;
    FFB_CUDA(cudaMemsetAsync(c.slab_counters, 0, 4 * sizeof(int), c.stream));
    if (c.n > 0) {
        k_route<<<(c.n + 255) / 256, 256, 0, c.stream>>>(streams_of(c, c.cur), streams_of(c, c.cur ^ 1), c.n, c.g.inv_dx,
                                                         k_begin, k_end, block_up, block_down, cap, c.slab_counters);
        launches++;
    }
    const int rows = slab_rows(c);
    if (block_up) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 1, cap, rows, block_up); launches++; }
    if (block_down) { k_write_header<<<1, 1, 0, c.stream>>>(c.slab_counters, 2, cap, rows, block_down); launches++; }
    FFB_CUDA(cudaMemcpyAsync(counts_host, c.slab_counters, 3 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    FFB_CUDA(cudaStreamSynchronize(c.stream));
    c.cur ^= 1;
    c.n = counts_host[0];
    c.sorted = false;
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
