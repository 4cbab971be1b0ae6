// ffb200_sort.cu -- cell binning, stable LSD radix sort on packed half-cell keys, bin table,
// SoA reorder. Integer work only; every result is bit-exact by construction.
//
// Replaces the reference's per-direction block binning (velocityadvector.cpp:190-465:
// _initializeBlockGrid, _computeGridCountData, the serial _sortParticlesIntoBlocks) with ONE
// sort per substep whose order serves P2G, G2P and advection:
//   key  = half-cell bin  (hx+A) + HX*((hy+A) + HZ-major...)  with h = floor(p * 2*(1/dx)), so
//          h >> 1 is exactly Grid3d::positionToGridIndex (grid3d.h:55-60);
//   ties = ascending original particle index, the order in which the reference accumulates a
//          block's particles (velocityadvector.cpp:383-413).
#include "ffb200_ctx.h"
#include "ffb200_seam.cuh"

#include <cstdlib>
#include <string>
#include <utility>

namespace ffb200 {

namespace {

constexpr int kSortThreads = 256;
#ifndef FFB_REORDER_THREADS
#define FFB_REORDER_THREADS 256
#endif
#ifndef FFB_KEYS_THREADS
#define FFB_KEYS_THREADS 256
#endif
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kKeysPerThread = 16;
constexpr int kTile = kSortThreads * kKeysPerThread;      // 4096 keys per CTA
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;      // 4096 entries per CTA

// ---- AoS <-> SoA --------------------------------------------------------------------------------

__global__ void k_unpack_aos(const float *__restrict__ aos, float *__restrict__ x, float *__restrict__ y,
                             float *__restrict__ z, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = aos[3 * (size_t)i + 0];
    y[i] = aos[3 * (size_t)i + 1];
    z[i] = aos[3 * (size_t)i + 2];
}

// out[orig[j]] = (x[j], y[j], z[j]): back to the host's particle order.
__global__ void k_pack_aos(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                           const uint32_t *__restrict__ orig, float *__restrict__ aos, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    size_t o = orig[j];
    aos[3 * o + 0] = x[j];
    aos[3 * o + 1] = y[j];
    aos[3 * o + 2] = z[j];
}

__global__ void k_iota(uint32_t *dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint32_t)i;
}

// ---- keys ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t half_cell_key(const GridDesc &g, float px, float py, float pz) {
    // floor(p * (2 * (1/dx))): doubling is exact in binary floating point, so (h >> 1) is
    // bit-for-bit Grid3d::positionToGridIndex(p, dx).
    const int hx = __double2int_rd((double)px * g.inv_2dx) + kApron;
    const int hy = __double2int_rd((double)py * g.inv_2dx) + kApron;
    const int hz = __double2int_rd((double)pz * g.inv_2dx) + kApron - 2 * g.kbase;
    if ((unsigned)hx >= (unsigned)g.HX || (unsigned)hy >= (unsigned)g.HY || (unsigned)hz >= (unsigned)g.HZ)
        return g.nbins;
    return (uint32_t)hx + (uint32_t)g.HX * ((uint32_t)hy + (uint32_t)g.HY * (uint32_t)hz);
}

// key[j], val[j] = j, and the bin histogram (integer atomics: the counts are order independent).
__global__ void k_keys(GridDesc g, const float *__restrict__ px, const float *__restrict__ py,
                       const float *__restrict__ pz, uint32_t *__restrict__ key, uint32_t *__restrict__ val,
                       uint32_t *__restrict__ bin_count, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t k = half_cell_key(g, px[j], py[j], pz[j]);
    key[j] = k;
    val[j] = (uint32_t)j;
    atomicAdd(bin_count + k, 1u);
}

// Counting-sort flavour: besides the key, keep the particle's arrival rank in its bin. Both kernels are bound by the
// latency of dependent scattered accesses (atomic with return value; bin_start[key] then a scattered store), at full
// occupancy already: each thread handles kSortIlp particles, a block stride apart, with the independent accesses of all
// of them issued before the first dependent one.
constexpr int kSortIlp = 4;

__global__ void k_keys_rank(GridDesc g, const float *__restrict__ px, const float *__restrict__ py,
                            const float *__restrict__ pz, uint32_t *__restrict__ key, uint32_t *__restrict__ rank,
                            uint32_t *__restrict__ bin_count, int n) {
    const int base = blockIdx.x * (blockDim.x * kSortIlp) + threadIdx.x;
    uint32_t k[kSortIlp], r[kSortIlp];
#pragma unroll
    for (int q = 0; q < kSortIlp; q++) {
        const int j = base + q * blockDim.x;
        if (j < n) k[q] = half_cell_key(g, px[j], py[j], pz[j]);
    }
#pragma unroll
    for (int q = 0; q < kSortIlp; q++)
        if (base + q * (int)blockDim.x < n) r[q] = atomicAdd(bin_count + k[q], 1u);
#pragma unroll
    for (int q = 0; q < kSortIlp; q++) {
        const int j = base + q * blockDim.x;
        if (j < n) {
            key[j] = k[q];
            rank[j] = r[q];
        }
    }
}

// slot = bin_start[key] + rank; sorted position -> (key, source slot).
__global__ void k_place(const uint32_t *__restrict__ key, const uint32_t *__restrict__ rank,
                        const uint32_t *__restrict__ bin_start, uint32_t *__restrict__ key_out,
                        uint32_t *__restrict__ val_out, int n) {
    const int base = blockIdx.x * (blockDim.x * kSortIlp) + threadIdx.x;
    uint32_t k[kSortIlp], dst[kSortIlp];
#pragma unroll
    for (int q = 0; q < kSortIlp; q++) {
        const int j = base + q * blockDim.x;
        if (j < n) {
            k[q] = key[j];
            dst[q] = rank[j];
        }
    }
#pragma unroll
    for (int q = 0; q < kSortIlp; q++)
        if (base + q * (int)blockDim.x < n) dst[q] += bin_start[k[q]];
#pragma unroll
    for (int q = 0; q < kSortIlp; q++) {
        const int j = base + q * blockDim.x;
        if (j < n) {
            key_out[dst[q]] = k[q];
            val_out[dst[q]] = (uint32_t)j;
        }
    }
}

// ---- exclusive scan (reduce / scan partials / apply) ---------------------------------------------------

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;                 // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;       // block total
    }
    __syncthreads();
    total = warp_sums[32];
    uint32_t r = warp_sums[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t *__restrict__ in, uint32_t *__restrict__ partial,
                                                               size_t n) {
    __shared__ uint32_t ws[33];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
    if (base + kScanItems <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) {
            uint4 t = p[q];
            s += t.x + t.y + t.z + t.w;
        }
    } else {
        for (int q = 0; q < kScanItems; q++)
            if (base + q < n) s += in[base + q];
    }
    uint32_t total;
    block_exclusive_scan(s, ws, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// one CTA, sequential over chunks of blockDim * kPartialItems (a 512^3 bin table has 262 K tile totals: 32 rounds)
constexpr int kPartialItems = 8;
__global__ void __launch_bounds__(1024) k_scan_partials(uint32_t *partial, int m) {
    __shared__ uint32_t ws[33];
    uint32_t carry = 0;
    for (int base = 0; base < m; base += blockDim.x * kPartialItems) {
        const int i0 = base + threadIdx.x * kPartialItems;
        uint32_t v[kPartialItems], s = 0;
#pragma unroll
        for (int q = 0; q < kPartialItems; q++) {
            v[q] = i0 + q < m ? partial[i0 + q] : 0u;
            s += v[q];
        }
        uint32_t total;
        uint32_t run = block_exclusive_scan(s, ws, total) + carry;
#pragma unroll
        for (int q = 0; q < kPartialItems; q++) {
            if (i0 + q < m) partial[i0 + q] = run;
            run += v[q];
        }
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(uint32_t *__restrict__ data, const uint32_t *__restrict__ partial,
                                                              size_t n) {
    __shared__ uint32_t ws[33];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
    const bool full = base + kScanItems <= n;
    if (full) {
        const uint4 *p = reinterpret_cast<const uint4 *>(data + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) {
            uint4 t = p[q];
            v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < kScanItems; q++) v[q] = (base + q < n) ? data[base + q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < kScanItems; q++) s += v[q];
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, ws, total) + partial[blockIdx.x];
#pragma unroll
    for (int q = 0; q < kScanItems; q++) {
        uint32_t t = v[q];
        v[q] = run;
        run += t;
    }
    if (full) {
        uint4 *p = reinterpret_cast<uint4 *>(data + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) p[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < kScanItems; q++)
            if (base + q < n) data[base + q] = v[q];
    }
}

int exclusive_scan_inplace(Context &c, uint32_t *data, size_t n) {
    if (n == 0) return 0;
    int blocks = (int)((n + kScanTile - 1) / kScanTile);
    if ((size_t)blocks > c.sort.scan_partials_cap) {
        if (c.sort.scan_partials) FFB_CUDA(cudaFree(c.sort.scan_partials));
        c.sort.scan_partials_cap = (size_t)blocks * 2;
        FFB_CUDA(cudaMalloc(&c.sort.scan_partials, c.sort.scan_partials_cap * sizeof(uint32_t)));
    }
    k_scan_reduce<<<blocks, kScanThreads, 0, c.stream>>>(data, c.sort.scan_partials, n);
    k_scan_partials<<<1, 1024, 0, c.stream>>>(c.sort.scan_partials, blocks);
    k_scan_apply<<<blocks, kScanThreads, 0, c.stream>>>(data, c.sort.scan_partials, n);
    return 3;
}

// ---- LSD radix sort pass -----------------------------------------------------------------------------
//
// A tile is 4096 consecutive keys; warp w of the CTA owns the 512 consecutive keys
// [w*512, (w+1)*512) and walks them in 16 rounds of 32 (coalesced). Ranking inside the tile is
// round-major per warp, warp-major per tile, i.e. ascending input position, so equal digits
// keep their input order: every pass is stable.

__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint32_t *__restrict__ key, uint32_t *__restrict__ tile_hist,
                                                              int n, int shift, int digits, int ntiles) {
    extern __shared__ uint32_t sh[];
    for (int d = threadIdx.x; d < digits; d += blockDim.x) sh[d] = 0;
    __syncthreads();
    const uint32_t mask = (uint32_t)digits - 1u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)warp * (32 * kKeysPerThread);
#pragma unroll 4
    for (int r = 0; r < kKeysPerThread; r++) {
        size_t i = base + (size_t)r * 32 + lane;
        if (i < (size_t)n) atomicAdd(&sh[(key[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < digits; d += blockDim.x) tile_hist[(size_t)d * ntiles + blockIdx.x] = sh[d];
}

__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const uint32_t *__restrict__ key_in, const uint32_t *__restrict__ val_in,
                                                                 uint32_t *__restrict__ key_out, uint32_t *__restrict__ val_out,
                                                                 const uint32_t *__restrict__ tile_off, int n, int shift,
                                                                 int digits, int ntiles) {
    extern __shared__ uint32_t wc[];                       // [kSortWarps][digits]
    for (int d = threadIdx.x; d < digits * kSortWarps; d += blockDim.x) wc[d] = 0;
    __syncthreads();
    const uint32_t mask = (uint32_t)digits - 1u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t *mywc = wc + warp * digits;
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)warp * (32 * kKeysPerThread);
    uint32_t k[kKeysPerThread], rank[kKeysPerThread];
#pragma unroll
    for (int r = 0; r < kKeysPerThread; r++) {
        size_t i = base + (size_t)r * 32 + lane;
        const bool ok = i < (size_t)n;
        k[r] = ok ? key_in[i] : 0u;
        const uint32_t dg = ok ? ((k[r] >> shift) & mask) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, dg);
        uint32_t cnt = 0;
        if (ok) cnt = mywc[dg];
        __syncwarp();
        if (ok && (peers & lt) == 0) mywc[dg] = cnt + __popc(peers);
        __syncwarp();
        rank[r] = cnt + __popc(peers & lt);
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps + this tile's global offset
    for (int d = threadIdx.x; d < digits; d += blockDim.x) {
        uint32_t run = tile_off[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            uint32_t t = wc[w * digits + d];
            wc[w * digits + d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kKeysPerThread; r++) {
        size_t i = base + (size_t)r * 32 + lane;
        if (i < (size_t)n) {
            uint32_t dst = mywc[(k[r] >> shift) & mask] + rank[r];
            key_out[dst] = k[r];
            val_out[dst] = val_in[i];
        }
    }
}

// ---- reorder -----------------------------------------------------------------------------------------
//
// Sorted position j takes the particle in old slot val[j]. The radix passes are stable with
// respect to the OLD slot order; when the old order is not the original order (resident
// multi-substep use) particles sharing a bin are re-ranked by their original index, so the
// result is always the (key, original index) order.
struct ReorderArgs {
    const float *src[15];
    float *dst[15];
    int nstreams;
};

template <bool SEAM>
__global__ void __launch_bounds__(FFB_REORDER_THREADS) k_reorder(const __grid_constant__ ReorderArgs a, const __grid_constant__ SeamParams sp,
                                                 const uint32_t *__restrict__ key, const uint32_t *__restrict__ val,
                                                 const uint32_t *__restrict__ bin_start, const uint32_t *__restrict__ orig_old,
                                                 uint32_t *__restrict__ orig_new, int n) {
    __shared__ int sh[SEAM ? 3 * FFB_REORDER_THREADS : 1];
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int hidx[3] = {-1, -1, -1};
    if (j < n) {
        const uint32_t slot = val[j];
        const uint32_t kk = key[j];
        const uint32_t s = bin_start[kk], e = bin_start[kk + 1];
        const uint32_t my = orig_old[slot];
        uint32_t dst = (uint32_t)j;
        if (e - s > 1u) {
            uint32_t rank = 0;
            for (uint32_t q = s; q < e; q++) rank += orig_old[val[q]] < my ? 1u : 0u;
            dst = s + rank;
        }
        orig_new[dst] = my;
        // all gathers first (independent, through the read-only path), then all stores: the stream
        // pointers may alias as far as the compiler knows, so a load/store-per-stream loop would
        // serialise on the memory latency
        float v[15];
#pragma unroll
        for (int t = 0; t < 15; t++)
            if (t < a.nstreams) v[t] = __ldg(a.src[t] + slot);
#pragma unroll
        for (int t = 0; t < 15; t++)
            if (t < a.nstreams) a.dst[t][dst] = v[t];
        // the P2G membership arithmetic of the particle rides along while the kernel waits for memory
        if (SEAM) seam_particle(sp, v[0], v[1], v[2], (int)dst, hidx);
    }
    if (SEAM) seam_mark_home(sp, hidx, sh);
}

__global__ void k_binning_dump(GridDesc g, const float *__restrict__ px, const float *__restrict__ py,
                               const float *__restrict__ pz, const uint32_t *__restrict__ orig,
                               int32_t *__restrict__ cell, uint32_t *__restrict__ hkey, uint32_t *__restrict__ perm, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float x = px[j], y = py[j], z = pz[j];
    const uint32_t o = orig[j];
    const int ci = pos2idx(x, g.inv_dx), cj = pos2idx(y, g.inv_dx), ck = pos2idx(z, g.inv_dx);
    cell[o] = in_range3(ci, cj, ck, g.I, g.J, g.K) ? ci + g.I * (cj + g.J * ck) : -1;
    hkey[o] = half_cell_key(g, x, y, z);
    perm[j] = o;
}

inline int blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

}  // namespace

int launch_exclusive_scan(Context &c, uint32_t *data, size_t n) { return exclusive_scan_inplace(c, data, n); }

int launch_unpack_aos(Context &c, const float *aos, float *const dst[3], int n) {
    if (n == 0) return 0;
    k_unpack_aos<<<blocks_for(n, 256), 256, 0, c.stream>>>(aos, dst[0], dst[1], dst[2], n);
    return 1;
}

int launch_pack_aos(Context &c, const float *const src[3], const uint32_t *orig, float *aos, int n) {
    if (n == 0) return 0;
    k_pack_aos<<<blocks_for(n, 256), 256, 0, c.stream>>>(src[0], src[1], src[2], orig, aos, n);
    return 1;
}

int launch_iota(Context &c, uint32_t *dst, int n) {
    if (n == 0) return 0;
    k_iota<<<blocks_for(n, 256), 256, 0, c.stream>>>(dst, n);
    return 1;
}

int launch_sort(Context &c, const SeamParams *seam) {
    const int n = c.n;
    int launches = 0;
    const GridDesc &g = c.g;
    ParticleSoA &src = c.soa[c.cur], &dst = c.soa[c.cur ^ 1];
    SortScratch &s = c.sort;

    // FFB200_SORT=radix selects the multi-pass LSD radix sort; the default is the single-pass
    // counting sort (radix = the whole key), possible because the dense bin table is needed anyway.
    static const bool use_radix = [] { const char *e = std::getenv("FFB200_SORT"); return e && std::string(e) == "radix"; }();

    // bin histogram + keys
    // (accumulating the scan's tile totals in the key pass -- one warp-aggregated atomic per tile -- to drop the scan's
    // reduce pass over the table was measured: +1.2 ms at 512^3, the match/atomic costs more than the 0.6 ms pass)
    FFB_CUDA(cudaMemsetAsync(s.bin_start, 0, ((size_t)g.nbins + 2) * sizeof(uint32_t), c.stream));
    if (n > 0) {
        if (use_radix)
            k_keys<<<blocks_for(n, 256), 256, 0, c.stream>>>(g, src.p[0], src.p[1], src.p[2], s.key[0], s.val[0], s.bin_start, n);
        else
            k_keys_rank<<<blocks_for(n, FFB_KEYS_THREADS * kSortIlp), FFB_KEYS_THREADS, 0, c.stream>>>(g, src.p[0], src.p[1], src.p[2], s.key[0], s.val[0], s.bin_start, n);
        launches++;
    }
    launches += exclusive_scan_inplace(c, s.bin_start, (size_t)g.nbins + 2);
    if (n == 0) {
        c.sorted = true;
        return launches;
    }

    if (!use_radix) {
        // slot = bin_start[key] + arrival rank. The arrival order inside a bin is arbitrary (integer
        // atomics); k_reorder re-ranks the members of every multi-particle bin by original index,
        // so the final order is the deterministic (key, original index) order all the same.
        k_place<<<blocks_for(n, FFB_KEYS_THREADS * kSortIlp), FFB_KEYS_THREADS, 0, c.stream>>>(s.key[0], s.val[0], s.bin_start, s.key[1], s.val[1], n);
        launches++;
        std::swap(s.key[0], s.key[1]);
        std::swap(s.val[0], s.val[1]);
    } else {
    // radix passes over the bits of [0, nbins]
    int bits = 1;
    while (bits < 32 && (g.nbins >> bits) != 0u) bits++;
    int passes = (bits + 7) / 8;
    if (passes > 3 && (bits + 2) / 3 <= 11) passes = 3;
    const int dbits = (bits + passes - 1) / passes;
    const int digits = 1 << dbits;
    const int ntiles = (n + kTile - 1) / kTile;
    const size_t need = (size_t)digits * ntiles;
    if (need > s.tile_hist_cap) {
        if (s.tile_hist) FFB_CUDA(cudaFree(s.tile_hist));
        s.tile_hist_cap = need + need / 4;
        FFB_CUDA(cudaMalloc(&s.tile_hist, s.tile_hist_cap * sizeof(uint32_t)));
    }
    const size_t smem_hist = (size_t)digits * sizeof(uint32_t);
    const size_t smem_scatter = (size_t)digits * kSortWarps * sizeof(uint32_t);
    if (smem_scatter > 48 * 1024)
        FFB_CUDA(cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
    int in = 0;
    for (int p = 0; p < passes; p++) {
        const int shift = p * dbits;
        k_radix_hist<<<ntiles, kSortThreads, smem_hist, c.stream>>>(s.key[in], s.tile_hist, n, shift, digits, ntiles);
        launches++;
        launches += exclusive_scan_inplace(c, s.tile_hist, need);
        k_radix_scatter<<<ntiles, kSortThreads, smem_scatter, c.stream>>>(s.key[in], s.val[in], s.key[in ^ 1], s.val[in ^ 1],
                                                                          s.tile_hist, n, shift, digits, ntiles);
        launches++;
        in ^= 1;
    }
    if (in != 0) {       // keep the sorted keys/vals in buffer 0 for later readers
        std::swap(s.key[0], s.key[1]);
        std::swap(s.val[0], s.val[1]);
    }
    }

    ReorderArgs a;
    int t = 0;
    for (int q = 0; q < 3; q++) { a.src[t] = src.p[q]; a.dst[t] = dst.p[q]; t++; }
    for (int q = 0; q < 3; q++) { a.src[t] = src.v[q]; a.dst[t] = dst.v[q]; t++; }
    if (c.has_affine)
        for (int q = 0; q < 9; q++) { a.src[t] = src.a[q]; a.dst[t] = dst.a[q]; t++; }
    a.nstreams = t;
    for (; t < 15; t++) { a.src[t] = nullptr; a.dst[t] = nullptr; }
    if (seam)
        k_reorder<true><<<blocks_for(n, FFB_REORDER_THREADS), FFB_REORDER_THREADS, 0, c.stream>>>(a, *seam, s.key[0], s.val[0], s.bin_start, src.orig, dst.orig, n);
    else
        k_reorder<false><<<blocks_for(n, FFB_REORDER_THREADS), FFB_REORDER_THREADS, 0, c.stream>>>(a, SeamParams{}, s.key[0], s.val[0], s.bin_start, src.orig, dst.orig, n);
    launches++;
    c.cur ^= 1;
    c.sorted = true;
    return launches;
}

int launch_binning_dump(Context &c, int32_t *cell, uint32_t *hkey, uint32_t *perm) {
    if (c.n == 0) return 0;
    ParticleSoA &s = c.soa[c.cur];
    k_binning_dump<<<blocks_for(c.n, 256), 256, 0, c.stream>>>(c.g, s.p[0], s.p[1], s.p[2], s.orig, cell, hkey, perm, c.n);
    return 1;
}

}  // namespace ffb200
